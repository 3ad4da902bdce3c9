// hm_plugin/TEncSAO_hevcdl.cpp -- drop-in definition of
//     Void TEncSampleAdaptiveOffset::getStatistics( SAOStatData*** blkStats, TComPicYuv* orgYuv, TComPicYuv* srcYuv, TComPic* pPic,
//                                                   Bool isCalculatePreDeblockSamples )
// (declared at HM_dl/source/Lib/TLibEncoder/TEncSampleAdaptiveOffset.h:114, reference body at TEncSampleAdaptiveOffset.cpp:
// 295-341; called from SAOProcess, :258) that computes the per-CTU SAO class statistics of the deblocked picture on the B200
// (hevcdl_sao_stats) when HEVCDL_SAO=1 and runs the reference's own pass otherwise.  The parameter decision that follows
// (decidePicParams / decideBlkParams: an RD search with CABAC bit estimates) is the reference's, on the host.
// Linked without editing the reference: sao_hook.h declares one extra member, getStatistics_reference, and hm_plugin/Makefile
// compiles the reference's TEncSampleAdaptiveOffset.cpp with its definition of getStatistics moved onto that name.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sao_hook.h"

#include "hevcdl.h"
#include "inloop_cache.h"

hevcdl_ctx *hevcdl_hm_context();                                     // TEncCu_hevcdl.cpp
void hevcdl_hm_count_sao( bool onDevice );
void hevcdl_hm_pin_picture( TComPicYuv *pic );                        // TEncCu_hevcdl.cpp

Void TEncSampleAdaptiveOffset::getStatistics( SAOStatData*** blkStats, TComPicYuv* orgYuv, TComPicYuv* srcYuv, TComPic* pPic, Bool isCalculatePreDeblockSamples )
{
  static const bool enabled = getenv( "HEVCDL_SAO" ) && atoi( getenv( "HEVCDL_SAO" ) ) == 1;
  hevcdl_ctx *ctx = enabled ? hevcdl_hm_context() : NULL;
  const TComSPS &sps = pPic->getPicSym()->getSPS();
  const TComPPS &pps = pPic->getPicSym()->getPPS();
  // what the device pass covers (csrc/sao.cuh): deblocked samples, skip lines of SAOLcuBoundary 0, 64x64 CTUs, 8-bit 4:2:0, one slice, no tiles
  const bool ok = ctx != NULL && !isCalculatePreDeblockSamples && m_maxCUWidth == 64 && m_maxCUHeight == 64 && m_chromaFormatIDC == CHROMA_420 &&
                  sps.getBitDepth( CHANNEL_TYPE_LUMA ) == 8 && sps.getBitDepth( CHANNEL_TYPE_CHROMA ) == 8 && pPic->getNumAllocatedSlice() == 1 &&
                  pps.getNumTileColumnsMinus1() == 0 && pps.getNumTileRowsMinus1() == 0 &&
                  m_skipLinesR[COMPONENT_Y][SAO_TYPE_EO_90] == 5 && m_skipLinesB[COMPONENT_Y][SAO_TYPE_BO] == 4 &&
                  m_skipLinesR[COMPONENT_Cb][SAO_TYPE_EO_90] == 3 && m_skipLinesB[COMPONENT_Cb][SAO_TYPE_BO] == 2;
  if ( !ok )
  {
    hevcdl_hm_count_sao( false );
    getStatistics_reference( blkStats, orgYuv, srcYuv, pPic, isCalculatePreDeblockSamples );
    return;
  }
  std::vector<int64_t> st( (size_t)m_numCTUsPic * 3 * NUM_SAO_NEW_TYPES * 2 * MAX_NUM_SAO_CLASSES );
  // taken already, in the deblocking call's round trip (TComLoopFilter_hevcdl.cpp), if srcYuv is still that deblocked picture
  const bool cached = g_hevcdl_inloop.valid && g_hevcdl_inloop.W == m_picWidth && g_hevcdl_inloop.H == m_picHeight && g_hevcdl_inloop.org == orgYuv &&
                      g_hevcdl_inloop.stats.size() == st.size() && hevcdl_inloop_guard( srcYuv ) == g_hevcdl_inloop.guard;
  if ( cached ) st = g_hevcdl_inloop.stats;
  if ( !cached ) { hevcdl_hm_pin_picture( orgYuv ); hevcdl_hm_pin_picture( srcYuv ); }
  const int rc = cached ? 0 : hevcdl_sao_stats( ctx, orgYuv->getAddr( COMPONENT_Y ), orgYuv->getAddr( COMPONENT_Cb ), orgYuv->getAddr( COMPONENT_Cr ),
                                   orgYuv->getStride( COMPONENT_Y ), orgYuv->getStride( COMPONENT_Cb ), srcYuv->getAddr( COMPONENT_Y ),
                                   srcYuv->getAddr( COMPONENT_Cb ), srcYuv->getAddr( COMPONENT_Cr ), srcYuv->getStride( COMPONENT_Y ),
                                   srcYuv->getStride( COMPONENT_Cb ), m_picWidth, m_picHeight, st.data() );
  if ( rc )
  {
    fprintf( stderr, "hevcdl: hevcdl_sao_stats failed: %s (%s)\n", hevcdl_status_str( rc ), hevcdl_last_error( ctx ) );
    exit( EXIT_FAILURE );
  }
  const int64_t *p = st.data();
  for ( Int a = 0; a < m_numCTUsPic; a++ )
    for ( Int c = 0; c < 3; c++ )
      for ( Int t = 0; t < NUM_SAO_NEW_TYPES; t++ )
      {
        for ( Int k = 0; k < MAX_NUM_SAO_CLASSES; k++ ) blkStats[a][c][t].diff[k] = *p++;
        for ( Int k = 0; k < MAX_NUM_SAO_CLASSES; k++ ) blkStats[a][c][t].count[k] = *p++;
      }
  if ( !cached ) g_hevcdl_inloop.valid = false;    // hevcdl_sao_stats reused the scratch the resident picture lived in
  hevcdl_hm_count_sao( true );
}
