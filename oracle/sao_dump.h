// oracle/sao_dump.h -- TEST INFRASTRUCTURE.  Included (by a sed-inserted #include) into a temp copy of the reference's
// TEncSampleAdaptiveOffset.cpp when oracle/Makefile builds the `TAppEncoder_saotrace` variant: an object declared at the top of
// TEncSampleAdaptiveOffset::getStatistics (HM_dl/source/Lib/TLibEncoder/TEncSampleAdaptiveOffset.cpp:295) writes, when the
// function returns, its inputs (original and deblocked pictures) and its output (the per-CTU / component / SAO type
// statistics) to the file named by $HEVCDL_SAO_DUMP.
// Record: int32 header[8] = {magic, W, H, numCTUs, preDeblock flag, 0, 0, 0}; int16 org Y,Cb,Cr; int16 src Y,Cb,Cr (dense);
//         then numCTUs x 3 components x 5 types x { int64 diff[32], int64 count[32] }.
#pragma once
#include <cstdio>
#include <cstdlib>

struct SaoDump
{
  SAOStatData ***st; TComPicYuv *org, *src; int nctu, pre;
  SaoDump( SAOStatData ***s, TComPicYuv *o, TComPicYuv *r, int n, bool p ) : st(s), org(o), src(r), nctu(n), pre(p) {}
  ~SaoDump()
  {
    static FILE *f = getenv("HEVCDL_SAO_DUMP") ? fopen(getenv("HEVCDL_SAO_DUMP"), "wb") : NULL;
    if (!f) return;
    int hdr[8] = { 0x53414F30, org->getWidth(COMPONENT_Y), org->getHeight(COMPONENT_Y), nctu, pre, 0, 0, 0 };
    fwrite(hdr, sizeof hdr, 1, f);
    TComPicYuv *pics[2] = { org, src };
    for (int k = 0; k < 2; k++)
      for (int c = 0; c < 3; c++)
      {
        const ComponentID id = ComponentID(c);
        const Pel *p = pics[k]->getAddr(id);
        for (int y = 0; y < pics[k]->getHeight(id); y++) fwrite(p + (size_t)y * pics[k]->getStride(id), sizeof(Pel), pics[k]->getWidth(id), f);
      }
    for (int a = 0; a < nctu; a++)
      for (int c = 0; c < 3; c++)
        for (int t = 0; t < NUM_SAO_NEW_TYPES; t++) { fwrite(st[a][c][t].diff, sizeof(Int64), MAX_NUM_SAO_CLASSES, f); fwrite(st[a][c][t].count, sizeof(Int64), MAX_NUM_SAO_CLASSES, f); }
    fflush(f);
  }
};
