// hm_plugin/TComSAO_hevcdl.cpp -- drop-in definition of
//     Void TComSampleAdaptiveOffset::offsetCTU( Int ctuRsAddr, TComPicYuv* srcYuv, TComPicYuv* resYuv, SAOBlkParam& saoblkParam, TComPic* pPic )
// (declared at HM_dl/source/Lib/TLibCommon/TComSampleAdaptiveOffset.h:81, reference body at TComSampleAdaptiveOffset.cpp:554-611;
// the encoder calls it once per CTU from decideBlkParams, TEncSampleAdaptiveOffset.cpp:894, right after the CTU's parameters are
// decided).  Every classification reads srcYuv (the deblocked copy), never resYuv, and nothing reads resYuv before the loop over
// the CTUs ends, so with HEVCDL_SAO=1 the calls only RECORD the resolved parameters and the call for the picture's last CTU
// applies all of them in one pass on the B200 (hevcdl_sao_apply).  Otherwise the reference's body runs.
// Linked without editing the reference: hm_plugin/Makefile compiles the reference's TComSampleAdaptiveOffset.cpp with
// -DoffsetCTU=offsetCTU_reference (definition and its decoder-side caller renamed together); here the class is read with that
// one extra member declared next to offsetCTU, so the fallback is a plain member call.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "TLibCommon/TComPic.h"
#define offsetCTU offsetCTU_reference( Int, TComPicYuv*, TComPicYuv*, SAOBlkParam&, TComPic* ); Void offsetCTU
#include "TLibCommon/TComSampleAdaptiveOffset.h"
#undef offsetCTU

#include "hevcdl.h"
#include "inloop_cache.h"

hevcdl_ctx *hevcdl_hm_context();                                     // TEncCu_hevcdl.cpp
void hevcdl_hm_count_sao_apply( bool onDevice );
void hevcdl_hm_pin_picture( TComPicYuv *pic );                        // TEncCu_hevcdl.cpp
void hevcdl_hm_count_inloop_resident( bool resident );

Void TComSampleAdaptiveOffset::offsetCTU( Int ctuRsAddr, TComPicYuv* srcYuv, TComPicYuv* resYuv, SAOBlkParam& saoblkParam, TComPic* pPic )
{
  static const bool enabled = getenv( "HEVCDL_SAO" ) && atoi( getenv( "HEVCDL_SAO" ) ) == 1;
  static std::vector<hevcdl_sao_param> prm;
  hevcdl_ctx *ctx = enabled ? hevcdl_hm_context() : NULL;
  const TComSPS &sps = pPic->getPicSym()->getSPS();
  const TComPPS &pps = pPic->getPicSym()->getPPS();
  // what the device pass covers (csrc/sao.cuh): 64x64 CTUs, 8-bit 4:2:0, one slice, no tiles
  const bool ok = ctx != NULL && m_maxCUWidth == 64 && m_maxCUHeight == 64 && m_chromaFormatIDC == CHROMA_420 &&
                  sps.getBitDepth( CHANNEL_TYPE_LUMA ) == 8 && sps.getBitDepth( CHANNEL_TYPE_CHROMA ) == 8 && pPic->getNumAllocatedSlice() == 1 &&
                  pps.getNumTileColumnsMinus1() == 0 && pps.getNumTileRowsMinus1() == 0 &&
                  ( ctuRsAddr == 0 || (Int)prm.size() == 3 * m_numCTUsPic );     // the picture's calls arrive in raster order from CTU 0
  if ( !ok )
  {
    if ( ctuRsAddr == m_numCTUsPic - 1 ) hevcdl_hm_count_sao_apply( false );
    offsetCTU_reference( ctuRsAddr, srcYuv, resYuv, saoblkParam, pPic );
    return;
  }
  if ( ctuRsAddr == 0 )
  {
    hevcdl_sao_param off;
    memset( &off, 0, sizeof off );
    off.type = -1;
    prm.assign( (size_t)3 * m_numCTUsPic, off );
  }
  for ( Int c = 0; c < 3; c++ )
  {
    const SAOOffset &o = saoblkParam[c];
    if ( o.modeIdc == SAO_MODE_OFF ) continue;
    hevcdl_sao_param &p = prm[(size_t)3 * ctuRsAddr + c];
    p.type = (int8_t)o.typeIdc;
    for ( Int k = 0; k < MAX_NUM_SAO_CLASSES; k++ ) p.offset[k] = (int8_t)o.offset[k];
  }
  if ( ctuRsAddr != m_numCTUsPic - 1 ) return;
  // the deblocked picture is still on the device if the fused in-loop call made it and srcYuv is that picture (inloop_cache.h)
  const bool resident = g_hevcdl_inloop.valid && g_hevcdl_inloop.W == m_picWidth && g_hevcdl_inloop.H == m_picHeight &&
                        hevcdl_inloop_guard( srcYuv ) == g_hevcdl_inloop.guard;
  g_hevcdl_inloop.valid = false;
  hevcdl_hm_pin_picture( resYuv );
  if ( !resident ) hevcdl_hm_pin_picture( srcYuv );
  const int rc = hevcdl_sao_apply( ctx, resident ? NULL : srcYuv->getAddr( COMPONENT_Y ), resident ? NULL : srcYuv->getAddr( COMPONENT_Cb ),
                                   resident ? NULL : srcYuv->getAddr( COMPONENT_Cr ),
                                   srcYuv->getStride( COMPONENT_Y ), srcYuv->getStride( COMPONENT_Cb ), resYuv->getAddr( COMPONENT_Y ),
                                   resYuv->getAddr( COMPONENT_Cb ), resYuv->getAddr( COMPONENT_Cr ), resYuv->getStride( COMPONENT_Y ),
                                   resYuv->getStride( COMPONENT_Cb ), m_picWidth, m_picHeight, prm.data() );
  if ( rc )
  {
    fprintf( stderr, "hevcdl: hevcdl_sao_apply failed: %s (%s)\n", hevcdl_status_str( rc ), hevcdl_last_error( ctx ) );
    exit( EXIT_FAILURE );
  }
  prm.clear();
  hevcdl_hm_count_sao_apply( true );
  hevcdl_hm_count_inloop_resident( resident );
}
