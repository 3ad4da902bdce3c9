// hm_plugin/TComPrediction_hevcdl.cpp -- drop-in definition of
//     Void TComPrediction::predIntraAng( const ComponentID compID, UInt uiDirMode, Pel* piOrg, UInt uiOrgStride, Pel* piPred,
//                                        UInt uiStride, TComTU &rTu, const Bool bUseFilteredPredSamples, const Bool bUseLosslessDPCM )
// (declared at HM_dl/source/Lib/TLibCommon/TComPrediction.h:117, reference body at TComPrediction.cpp:390-472; called by the RD
// pass for every luma / chroma transform block, TEncSearch.cpp:1208, and by the first pass for every mode, :2303) that computes
// the block on the B200 (hevcdl_intra_pred) from HM's own reference samples when HEVCDL_PRED=1 and runs the reference's body
// otherwise.  Linked without editing the reference: hm_plugin/Makefile compiles the reference's TComPrediction.cpp with
// -DpredIntraAng=predIntraAng_reference; ref_pred_call.cpp (same rename) is the trampoline back to that body.
#include <cstdio>
#include <cstdlib>

#include "TLibCommon/TComPrediction.h"
#include "TLibCommon/TComTU.h"
#include "TLibCommon/TComDataCU.h"
#include "TLibCommon/TComPic.h"

#include "hevcdl.h"

hevcdl_ctx *hevcdl_hm_context();                                     // TEncCu_hevcdl.cpp
void hevcdl_hm_count_pred( bool onDevice );
void hevcdl_ref_predIntraAng( TComPrediction *p, const ComponentID compID, UInt uiDirMode, Pel *piOrg, UInt uiOrgStride, Pel *piPred, UInt uiStride,
                              TComTU &rTu, const Bool bUseFilteredPredSamples, const Bool bUseLosslessDPCM );   // ref_pred_call.cpp

Void TComPrediction::predIntraAng( const ComponentID compID, UInt uiDirMode, Pel* piOrg, UInt uiOrgStride, Pel* piPred, UInt uiStride, TComTU &rTu,
                                   const Bool bUseFilteredPredSamples, const Bool bUseLosslessDPCM )
{
  static const bool enabled = getenv( "HEVCDL_PRED" ) && atoi( getenv( "HEVCDL_PRED" ) ) == 1;
  hevcdl_ctx *ctx = enabled ? hevcdl_hm_context() : NULL;
  const TComRectangle &rect = rTu.getRect( isLuma( compID ) ? COMPONENT_Y : COMPONENT_Cb );
  const Int n = rect.width;
  TComDataCU *pcCU = rTu.getCU();
  // what the device predictor covers (csrc/pred.cuh): square 4..64 blocks, 8-bit samples, no lossless DPCM
  const bool ok = ctx != NULL && !bUseLosslessDPCM && (Int)rect.height == n && n >= 4 && n <= 64 && ( n & ( n - 1 ) ) == 0 && uiDirMode < 35 &&
                  pcCU->getSlice()->getSPS()->getBitDepth( toChannelType( compID ) ) == 8;
  if ( !ok )
  {
    hevcdl_hm_count_pred( false );
    hevcdl_ref_predIntraAng( this, compID, uiDirMode, piOrg, uiOrgStride, piPred, uiStride, rTu, bUseFilteredPredSamples, bUseLosslessDPCM );
    return;
  }
  const Pel *src = getPredictorPtr( compID, bUseFilteredPredSamples );   // (2n+1) x (2n+1), first row and first column used
  const Int sw = 2 * n + 1;
  int16_t line[4 * 64 + 1], blk[64 * 64];
  for ( Int k = 0; k < 2 * n; k++ ) line[k] = src[( 2 * n - k ) * sw];   // left column, bottom-up
  for ( Int k = 0; k <= 2 * n; k++ ) line[2 * n + k] = src[k];           // corner, top row
  const UInt uiAbsPartIdx = rTu.GetAbsPartIdxTU();
  const bool edge = isLuma( compID ) && !( pcCU->isRDPCMEnabled( uiAbsPartIdx ) && pcCU->getCUTransquantBypass( uiAbsPartIdx ) );
  Int lg = 0;
  while ( ( 1 << lg ) < n ) lg++;
  hevcdl_pred_req rq = { (uint8_t)lg, (uint8_t)uiDirMode, (uint8_t)( edge ? HEVCDL_PRED_EDGE : 0 ), 0, 0, 0 };
  const int rc = hevcdl_intra_pred( ctx, 1, &rq, line, (size_t)( 4 * n + 1 ), blk, (size_t)n * n );
  if ( rc )
  {
    fprintf( stderr, "hevcdl: hevcdl_intra_pred failed: %s (%s)\n", hevcdl_status_str( rc ), hevcdl_last_error( ctx ) );
    exit( EXIT_FAILURE );
  }
  for ( Int y = 0; y < n; y++ )
    for ( Int x = 0; x < n; x++ ) piPred[y * uiStride + x] = blk[y * n + x];
  hevcdl_hm_count_pred( true );
}
