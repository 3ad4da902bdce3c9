// hm_plugin/TComLoopFilter_hevcdl.cpp -- drop-in definition of
//     Void TComLoopFilter::loopFilterPic( TComPic* pcPic )
// (declared at HM_dl/source/Lib/TLibCommon/TComLoopFilter.h, reference body at TComLoopFilter.cpp:130-158; called once per
// picture from TEncGOP::compressGOP, TEncGOP.cpp:1742) that runs the deblocking filter of an all-intra picture on the B200
// (hevcdl_deblock_frame) when HEVCDL_DBF=1, and the reference's own filter otherwise.
//
// Linked without editing the reference, like compressCtu: hm_plugin/Makefile compiles the reference's TComLoopFilter.cpp with
// -DloopFilterPic=loopFilterPic_reference (its body keeps its code under another name) and this translation unit provides the
// symbol TEncGOP calls; ref_loopfilter_call.cpp, compiled with the same rename, is the trampoline back to the reference body.
// What the device filter needs from the coded picture is what the reference's filter reads from it: per 4x4 luma unit the
// size of the transform unit covering it (CU depth + transform index) and its QP, the slice's deblocking offsets and the PPS
// chroma QP offsets.  The picture must be one the device filter covers (every CU intra -- the all-intra configurations of
// the reference --, one slice, no tiles, no PCM / lossless blocks, 8-bit 4:2:0, no deblocking metric); anything else falls
// back to the reference's filter, which is the same arithmetic on the host.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "TLibCommon/TComLoopFilter.h"
#include "TLibCommon/TComPic.h"

#include "hevcdl.h"

#include "inloop_cache.h"

hevcdl_ctx *hevcdl_hm_context();                                     // TEncCu_hevcdl.cpp: the encoder's session (NULL before the first CTU)
void hevcdl_hm_count_dbf( bool onDevice );
void hevcdl_hm_pin_picture( TComPicYuv *pic );                        // TEncCu_hevcdl.cpp
HevcdlInloopCache g_hevcdl_inloop;                                   // inloop_cache.h
void hevcdl_ref_loopFilterPic( TComLoopFilter *lf, TComPic *pcPic ); // ref_loopfilter_call.cpp

Void TComLoopFilter::loopFilterPic( TComPic* pcPic )
{
  static const bool enabled = getenv( "HEVCDL_DBF" ) && atoi( getenv( "HEVCDL_DBF" ) ) == 1;
  hevcdl_ctx *ctx = enabled ? hevcdl_hm_context() : NULL;
  TComSlice *sl = pcPic->getSlice( 0 );
  const TComSPS *sps = sl->getSPS();
  const TComPPS *pps = sl->getPPS();
  TComPicYuv *rec = pcPic->getPicYuvRec();
  bool ok = ctx != NULL && pcPic->getNumAllocatedSlice() == 1 && !sl->getDeblockingFilterDisable() && rec->getChromaFormat() == CHROMA_420 &&
            sps->getBitDepth( CHANNEL_TYPE_LUMA ) == 8 && sps->getBitDepth( CHANNEL_TYPE_CHROMA ) == 8 && !sps->getUsePCM() &&
            !pps->getTransquantBypassEnabledFlag() && pps->getNumTileColumnsMinus1() == 0 && pps->getNumTileRowsMinus1() == 0 &&
            sps->getMaxCUWidth() == 64 && sps->getMaxCUHeight() == 64;
  const int W = rec->getWidth( COMPONENT_Y ), H = rec->getHeight( COMPONENT_Y );
  std::vector<uint8_t> tu;
  std::vector<int8_t> qp;
  if ( ok )
  {
    const int w4 = W / 4, h4 = H / 4;
    tu.assign( (size_t)w4 * h4, 0 );
    qp.assign( (size_t)w4 * h4, 0 );
    const UInt ctuW = pcPic->getFrameWidthInCtus(), nPart = pcPic->getNumPartitionsInCtu(), partW = pcPic->getNumPartInCtuWidth();
    for ( UInt a = 0; a < pcPic->getNumberOfCtusInFrame() && ok; a++ )
    {
      TComDataCU *c = pcPic->getCtu( a );
      for ( UInt z = 0; z < nPart; z++ )
      {
        const UInt r = g_auiZscanToRaster[z];
        const int x = ( a % ctuW ) * 64 + ( r % partW ) * 4, y = ( a / ctuW ) * 64 + ( r / partW ) * 4;
        if ( x >= W || y >= H ) continue;
        if ( !c->isIntra( z ) || c->getQP( z ) < 0 || c->getQP( z ) > 51 ) { ok = false; break; }
        const int lg = 6 - (int)c->getDepth( z ) - (int)c->getTransformIdx( z );
        tu[(size_t)( y / 4 ) * w4 + x / 4] = (uint8_t)( lg < 2 ? 2 : ( lg > 5 ? 5 : lg ) );
        qp[(size_t)( y / 4 ) * w4 + x / 4] = (int8_t)c->getQP( z );
      }
    }
  }
  if ( !ok )
  {
    hevcdl_hm_count_dbf( false );
    hevcdl_ref_loopFilterPic( this, pcPic );
    return;
  }
  // With HEVCDL_SAO=1 as well, and SAO enabled for the sequence, the statistics SAOProcess will ask for next are taken in the same
  // round trip (hevcdl_inloop_frame) and the deblocked picture stays on the device for the offset pass: TEncSAO_hevcdl.cpp and
  // TComSAO_hevcdl.cpp pick both up from g_hevcdl_inloop after checking that the picture they are handed is still this one.
  hevcdl_hm_pin_picture( rec );
  static const bool fuse = getenv( "HEVCDL_SAO" ) && atoi( getenv( "HEVCDL_SAO" ) ) == 1 && !( getenv( "HEVCDL_INLOOP_FUSE" ) && atoi( getenv( "HEVCDL_INLOOP_FUSE" ) ) == 0 );
  g_hevcdl_inloop.valid = false;
  int rc;
  if ( fuse && sps->getUseSAO() )
  {
    TComPicYuv *org = pcPic->getPicYuvOrg();
    hevcdl_hm_pin_picture( org );
    const size_t nst = (size_t)( ( W + 63 ) / 64 ) * ( ( H + 63 ) / 64 ) * 3 * 5 * 64;
    g_hevcdl_inloop.stats.resize( nst );
    rc = hevcdl_inloop_frame( ctx, rec->getAddr( COMPONENT_Y ), rec->getStride( COMPONENT_Y ), rec->getAddr( COMPONENT_Cb ),
                              rec->getAddr( COMPONENT_Cr ), rec->getStride( COMPONENT_Cb ), W, H, tu.data(), qp.data(),
                              sl->getDeblockingFilterBetaOffsetDiv2(), sl->getDeblockingFilterTcOffsetDiv2(),
                              pps->getQpOffset( COMPONENT_Cb ), pps->getQpOffset( COMPONENT_Cr ), org->getAddr( COMPONENT_Y ),
                              org->getAddr( COMPONENT_Cb ), org->getAddr( COMPONENT_Cr ), org->getStride( COMPONENT_Y ),
                              org->getStride( COMPONENT_Cb ), g_hevcdl_inloop.stats.data() );
    if ( rc == 0 )
    {
      g_hevcdl_inloop.valid = true;
      g_hevcdl_inloop.W = W; g_hevcdl_inloop.H = H;
      g_hevcdl_inloop.guard = hevcdl_inloop_guard( rec );
      g_hevcdl_inloop.org = org;
    }
  }
  else
  {
    rc = hevcdl_deblock_frame( ctx, rec->getAddr( COMPONENT_Y ), rec->getStride( COMPONENT_Y ), rec->getAddr( COMPONENT_Cb ),
                               rec->getAddr( COMPONENT_Cr ), rec->getStride( COMPONENT_Cb ), W, H, tu.data(), qp.data(),
                               sl->getDeblockingFilterBetaOffsetDiv2(), sl->getDeblockingFilterTcOffsetDiv2(),
                               pps->getQpOffset( COMPONENT_Cb ), pps->getQpOffset( COMPONENT_Cr ) );
  }
  if ( rc )
  {
    fprintf( stderr, "hevcdl: hevcdl_deblock_frame / hevcdl_inloop_frame failed: %s (%s)\n", hevcdl_status_str( rc ), hevcdl_last_error( ctx ) );
    exit( EXIT_FAILURE );
  }
  hevcdl_hm_count_dbf( true );
}
