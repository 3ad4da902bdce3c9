// oracle/pred_dump.h -- TEST INFRASTRUCTURE.  Included (by a sed-inserted #include) into a temp copy of the reference's
// TComPrediction.cpp when oracle/Makefile builds the `TAppEncoder_predtrace` variant: an object declared at the top of
// TComPrediction::predIntraAng (HM_dl/source/Lib/TLibCommon/TComPrediction.cpp:390) writes, when the function returns, the
// reference samples it read (getPredictorPtr: first row and first column of the (2N+1)^2 array) and the block it predicted
// to the file named by $HEVCDL_PRED_DUMP -- at most 6 calls per (luma / chroma, block size, mode).
// Record: int32 header[8] = {magic, component, mode, N, filtered-references flag, edge filters enabled, 0, 0};
//         int16 line[4N+1] (left column bottom-up, corner, top row left to right); int16 pred[N*N].
#pragma once
#include <cstdio>
#include <cstdlib>

struct PredDump
{
  TComPrediction *self; ComponentID comp; UInt mode; Pel *pred; UInt stride; TComTU &tu; Bool filt, dpcm;
  PredDump( TComPrediction *s, ComponentID c, UInt m, Pel *p, UInt st, TComTU &t, Bool f, Bool d ) : self(s), comp(c), mode(m), pred(p), stride(st), tu(t), filt(f), dpcm(d) {}
  ~PredDump()
  {
    static FILE *f = getenv("HEVCDL_PRED_DUMP") ? fopen(getenv("HEVCDL_PRED_DUMP"), "wb") : NULL;
    static unsigned char seen[2][6][35];
    if (!f || dpcm) return;
    const TComRectangle &rect = tu.getRect(isLuma(comp) ? COMPONENT_Y : COMPONENT_Cb);
    const int n = rect.width;
    if (rect.height != (UInt)n) return;
    int lg = 0; while ((1 << lg) < n) lg++;
    unsigned char &cnt = seen[isLuma(comp) ? 0 : 1][lg][mode];
    if (cnt >= 6) return;
    cnt++;
    TComDataCU *cu = tu.getCU();
    const UInt idx = tu.GetAbsPartIdxTU();
    const int edge = !(cu->isRDPCMEnabled(idx) && cu->getCUTransquantBypass(idx));
    int hdr[8] = { 0x50524430, (int)comp, (int)mode, n, filt ? 1 : 0, edge, 0, 0 };
    fwrite(hdr, sizeof hdr, 1, f);
    const Pel *src = self->getPredictorPtr(comp, filt);
    const int sw = 2 * n + 1;
    for (int k = 2 * n; k >= 1; k--) fwrite(src + k * sw, sizeof(Pel), 1, f);    // left column, bottom-up
    fwrite(src, sizeof(Pel), 2 * n + 1, f);                                       // corner + top row
    for (int y = 0; y < n; y++) fwrite(pred + (size_t)y * stride, sizeof(Pel), n, f);
    fflush(f);
  }
};
