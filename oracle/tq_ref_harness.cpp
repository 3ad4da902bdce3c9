// oracle/tq_ref_harness.cpp -- TEST INFRASTRUCTURE.  Thin C entry points onto the REFERENCE's own 2-D core transforms
// (free functions xTrMxN / xITrMxN, HM_dl/source/Lib/TLibCommon/TComTrQuant.cpp:860,927), linked from oracle/_ref/libhmref.a
// (the reference compiled unmodified by oracle/Makefile).  tools/gen_golden_tq.py feeds them random blocks to make
// tests/golden/tq_transform_ref.npz, the pin of oracle/tq_oracle.c's transforms.  Built only where /root/reference exists.
#include "TLibCommon/TComRom.h"
#include "TLibCommon/TypeDef.h"

Void xTrMxN(Int bitDepth, TCoeff *block, TCoeff *coeff, Int iWidth, Int iHeight, Bool useDST, const Int maxLog2TrDynamicRange);
Void xITrMxN(Int bitDepth, TCoeff *coeff, TCoeff *block, Int iWidth, Int iHeight, Bool useDST, const Int maxLog2TrDynamicRange);

extern "C" {
void tqref_init() { initROM(); }
void tqref_forward(int *block, int *coeff, int n, int use_dst) { xTrMxN(8, block, coeff, n, n, use_dst != 0, 15); }
void tqref_inverse(int *coeff, int *block, int n, int use_dst) { xITrMxN(8, coeff, block, n, n, use_dst != 0, 15); }
}
