// oracle/saoapply_dump.h -- TEST INFRASTRUCTURE.  Included (by a sed-inserted #include) into a temp copy of the reference's
// TComSampleAdaptiveOffset.cpp when oracle/Makefile builds the `TAppEncoder_saoapplytrace` variant: an object declared at the
// top of TComSampleAdaptiveOffset::offsetCTU (HM_dl/source/Lib/TLibCommon/TComSampleAdaptiveOffset.cpp:554) collects the
// resolved SAO parameters of every CTU and, when the call for the picture's last CTU returns, writes the deblocked picture
// (srcYuv), the parameters and the picture with the offsets applied (resYuv) to the file named by $HEVCDL_SAOAPPLY_DUMP.
// Record: int32 header[8] = {magic, W, H, numCTUs, 0, 0, 0, 0}; int16 src Y,Cb,Cr (dense); int8 type[numCTUs*3] (-1 = off);
//         int8 offset[numCTUs*3][32]; int16 res Y,Cb,Cr (dense).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>

struct SaoApplyDump
{
  int ctu, nctu; TComPicYuv *src, *res; SAOBlkParam &prm;
  static std::vector<signed char> &types() { static std::vector<signed char> v; return v; }
  static std::vector<signed char> &offs() { static std::vector<signed char> v; return v; }
  SaoApplyDump( int c, int n, TComPicYuv *s, TComPicYuv *r, SAOBlkParam &p ) : ctu(c), nctu(n), src(s), res(r), prm(p)
  {
    if (ctu == 0) { types().assign((size_t)nctu * 3, -1); offs().assign((size_t)nctu * 3 * 32, 0); }
    if ((int)types().size() != nctu * 3) return;
    for (int c3 = 0; c3 < 3; c3++)
      if (prm[c3].modeIdc != SAO_MODE_OFF)
      {
        types()[ctu * 3 + c3] = (signed char)prm[c3].typeIdc;
        for (int k = 0; k < 32; k++) offs()[((size_t)ctu * 3 + c3) * 32 + k] = (signed char)prm[c3].offset[k];
      }
  }
  ~SaoApplyDump()
  {
    static FILE *f = getenv("HEVCDL_SAOAPPLY_DUMP") ? fopen(getenv("HEVCDL_SAOAPPLY_DUMP"), "wb") : NULL;
    if (!f || ctu != nctu - 1 || (int)types().size() != nctu * 3) return;
    int hdr[8] = { 0x53414F41, src->getWidth(COMPONENT_Y), src->getHeight(COMPONENT_Y), nctu, 0, 0, 0, 0 };
    fwrite(hdr, sizeof hdr, 1, f);
    for (int k = 0; k < 2; k++)
    {
      TComPicYuv *pic = k ? res : src;
      for (int c = 0; c < 3; c++)
      {
        const ComponentID id = ComponentID(c);
        const Pel *p = pic->getAddr(id);
        for (int y = 0; y < pic->getHeight(id); y++) fwrite(p + (size_t)y * pic->getStride(id), sizeof(Pel), pic->getWidth(id), f);
      }
      if (k == 0) { fwrite(types().data(), 1, types().size(), f); fwrite(offs().data(), 1, offs().size(), f); }
    }
    fflush(f);
  }
};
