// oracle/rdoq_dump.h -- TEST INFRASTRUCTURE.  Included (by a sed-inserted #include) into a temp copy of the reference's
// TComTrQuant.cpp when oracle/Makefile builds the `TAppEncoder_rdoqtrace` variant: an object declared at the top of
// TComTrQuant::xRateDistOptQuant (HM_dl/source/Lib/TLibCommon/TComTrQuant.cpp:2119) records the call's inputs and, when the
// function returns, its outputs into the binary file named by $HEVCDL_RDOQ_DUMP.  tools/gen_golden_tq.py turns a sample of
// the records into tests/golden/tq_rdoq_192x128_qp32.npz, the pin of oracle/rdoq_oracle.c and of the device RDOQ.
// Record: int32 header[16] = {magic, width, compID, qp, per, rem, scanType, transformSkip, ctxCbf (offset included), isIntra,
//         sdh, golombInit, sizeof(estBitsSbacStruct), trIdxIsZero, 0, 0}; double lambda, errScale; estBitsSbacStruct;
//         int32 src[width^2]; int32 dst[width^2]; int32 absSum.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>

struct RdoqDump
{
  FILE *f;
  const TCoeff *dst; const TCoeff *absSum; int n2;
  RdoqDump( TComTU &rTu, const TCoeff *src, const TCoeff *dst_, const TCoeff &absSum_, const ComponentID compID, const QpParam &cQP,
            double lambda, const estBitsSbacStruct *est, double errScale ) : f(NULL), dst(dst_), absSum(&absSum_), n2(0)
  {
    static FILE *file = getenv("HEVCDL_RDOQ_DUMP") ? fopen(getenv("HEVCDL_RDOQ_DUMP"), "wb") : NULL;
    if (!file) return;
    f = file;
    TComDataCU *cu = rTu.getCU();
    const UInt part = rTu.GetAbsPartIdxTU();
    const TComRectangle &rect = rTu.getRect(compID);
    TUEntropyCodingParameters cp;
    getTUEntropyCodingParameters(cp, rTu, compID);
    n2 = rect.width * rect.height;
    int hdr[16] = { 0x52444F51, (int)rect.width, (int)compID, cQP.Qp, cQP.per, cQP.rem, (int)cp.scanType, (int)cu->getTransformSkip(part, compID),
                    (int)(cu->getCtxQtCbf(rTu, toChannelType(compID)) + getCBFContextOffset(compID)), (int)cu->isIntra(part),
                    (int)cu->getSlice()->getPPS()->getSignDataHidingEnabledFlag(),
                    (int)(est->golombRiceAdaptationStatistics[rTu.getGolombRiceStatisticsIndex(compID)] / RExt__GOLOMB_RICE_INCREMENT_DIVISOR),
                    (int)sizeof(estBitsSbacStruct), (int)(cu->getTransformIdx(part) == 0), 0, 0 };
    fwrite(hdr, sizeof hdr, 1, f);
    fwrite(&lambda, sizeof(double), 1, f);
    fwrite(&errScale, sizeof(double), 1, f);
    fwrite(est, sizeof(estBitsSbacStruct), 1, f);
    fwrite(src, sizeof(TCoeff), n2, f);
  }
  ~RdoqDump()
  {
    if (!f) return;
    fwrite(dst, sizeof(TCoeff), n2, f);
    int a = (int)*absSum;
    fwrite(&a, sizeof(int), 1, f);
    fflush(f);
  }
};
