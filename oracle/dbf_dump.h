// oracle/dbf_dump.h -- TEST INFRASTRUCTURE.  Included (by a sed-inserted #include) into a temp copy of the reference's
// TEncGOP.cpp when oracle/Makefile builds the `TAppEncoder_dbftrace` variant: the reconstructed picture is written to the file
// named by $HEVCDL_DBF_DUMP right before and right after m_pcLoopFilter->loopFilterPic( pcPic ) (TEncGOP.cpp:1742), together with
// what the deblocking filter reads from the coded picture: per 4x4 luma unit the log2 size of the transform unit covering it
// and its QP, plus the slice / PPS offsets.  tools/gen_golden_tq.py turns the records into tests/golden/dbf_*.npz.
// Record: int32 header[12] = {magic, phase (0 before / 1 after), W, H, betaOffsetDiv2, tcOffsetDiv2, cbQpOffset, crQpOffset,
//         deblockingDisabled, allIntra, 0, 0}; int16 Y[H*W], Cb[(H/2)*(W/2)], Cr[...]; then (phase 0 only) uint8 tuLog2[(H/4)*(W/4)],
//         int8 qp[(H/4)*(W/4)].
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>

static inline void hevcdlDbfDump( TComPic *pic, int phase )
{
  static FILE *f = getenv("HEVCDL_DBF_DUMP") ? fopen(getenv("HEVCDL_DBF_DUMP"), "wb") : NULL;
  if (!f) return;
  TComPicYuv *rec = pic->getPicYuvRec();
  const int W = rec->getWidth(COMPONENT_Y), H = rec->getHeight(COMPONENT_Y);
  TComSlice *sl = pic->getSlice(0);
  const int w4 = W / 4, h4 = H / 4;
  std::vector<unsigned char> tu(w4 * h4, 0);
  std::vector<signed char> qp(w4 * h4, 0);
  int allIntra = 1;
  const UInt ctuW = pic->getFrameWidthInCtus(), maxCU = sl->getSPS()->getMaxCUWidth(), nPart = pic->getNumPartitionsInCtu(), partW = pic->getNumPartInCtuWidth();
  for (UInt a = 0; a < pic->getNumberOfCtusInFrame(); a++)
  {
    TComDataCU *c = pic->getCtu(a);
    for (UInt z = 0; z < nPart; z++)
    {
      const UInt r = g_auiZscanToRaster[z];
      const int x = (a % ctuW) * maxCU + (r % partW) * 4, y = (a / ctuW) * maxCU + (r / partW) * 4;
      if (x >= W || y >= H) continue;
      int lg = 6 - (int)c->getDepth(z) - (int)c->getTransformIdx(z);
      tu[(y / 4) * w4 + x / 4] = (unsigned char)(lg < 2 ? 2 : lg);
      qp[(y / 4) * w4 + x / 4] = (signed char)c->getQP(z);
      if (!c->isIntra(z)) allIntra = 0;
    }
  }
  int hdr[12] = { 0x44424630, phase, W, H, sl->getDeblockingFilterBetaOffsetDiv2(), sl->getDeblockingFilterTcOffsetDiv2(),
                  sl->getPPS()->getQpOffset(COMPONENT_Cb), sl->getPPS()->getQpOffset(COMPONENT_Cr), (int)sl->getDeblockingFilterDisable(), allIntra, 0, 0 };
  fwrite(hdr, sizeof hdr, 1, f);
  for (int comp = 0; comp < 3; comp++)
  {
    const ComponentID id = ComponentID(comp);
    const Pel *p = rec->getAddr(id);
    for (int y = 0; y < rec->getHeight(id); y++) fwrite(p + (size_t)y * rec->getStride(id), sizeof(Pel), rec->getWidth(id), f);
  }
  if (phase == 0) { fwrite(tu.data(), 1, tu.size(), f); fwrite(qp.data(), 1, qp.size(), f); }
  fflush(f);
}
