// hm_plugin/inloop_cache.h -- what the fused in-loop call (hevcdl_inloop_frame, made by TComLoopFilter_hevcdl.cpp when HEVCDL_DBF=1
// and HEVCDL_SAO=1) leaves for the two SAO hooks of the same picture: the statistics (consumed by TEncSAO_hevcdl.cpp instead of a
// second upload of both pictures) and the fact that the deblocked picture is resident on the device (TComSAO_hevcdl.cpp then
// applies the offsets with src = NULL).  Both hooks first check that the picture HM hands them IS that deblocked picture: same
// size and the same hash over every sample of 16 luma rows and 8 rows of each chroma plane -- HM copies the reconstruction into
// its SAO source buffer unchanged (TEncSampleAdaptiveOffset.cpp:253-256), but the hooks do not take that on trust.
#ifndef HEVCDL_INLOOP_CACHE_H
#define HEVCDL_INLOOP_CACHE_H
#include <cstdint>
#include <vector>

#include "TLibCommon/TComPicYuv.h"

struct HevcdlInloopCache
{
  bool valid = false;
  int W = 0, H = 0;
  uint64_t guard = 0;
  const TComPicYuv *org = NULL;
  std::vector<int64_t> stats;      // [nctu][3][5][2][32]
};
extern HevcdlInloopCache g_hevcdl_inloop;   // TComLoopFilter_hevcdl.cpp

static inline uint64_t hevcdl_inloop_guard( TComPicYuv *pic )
{
  uint64_t h = 1469598103934665603ull;
  for ( int c = 0; c < 3; c++ )
  {
    const ComponentID id = ComponentID( c );
    const int w = pic->getWidth( id ), hgt = pic->getHeight( id ), rows = c ? 8 : 16;
    const Pel *p = pic->getAddr( id );
    for ( int k = 0; k < rows; k++ )
    {
      const Pel *r = p + (size_t)( (long long)k * ( hgt - 1 ) / ( rows - 1 ) ) * pic->getStride( id );
      for ( int x = 0; x < w; x++ ) { h ^= (uint64_t)(uint16_t)r[x]; h *= 1099511628211ull; }
    }
  }
  return h;
}
#endif
