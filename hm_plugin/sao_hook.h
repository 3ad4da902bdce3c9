// hm_plugin/sao_hook.h -- reads the reference's TEncSampleAdaptiveOffset class (TLibEncoder/TEncSampleAdaptiveOffset.h)
// with ONE extra private member declared next to getStatistics (:114):
//     Void getStatistics_reference( SAOStatData***, TComPicYuv*, TComPicYuv*, TComPic*, Bool );
// The Makefile compiles a copy of the reference's TEncSampleAdaptiveOffset.cpp in which only the DEFINITION line (:295) carries
// that name (one sed substitution; the callers at :258 and :278 keep calling getStatistics), and TEncSAO_hevcdl.cpp defines
// getStatistics itself: device statistics when HEVCDL_SAO=1, getStatistics_reference otherwise.  The reference tree is not edited.
#ifndef HEVCDL_SAO_HOOK_H
#define HEVCDL_SAO_HOOK_H
#include "TLibCommon/TComSampleAdaptiveOffset.h"
#include "TLibCommon/TComPic.h"
#include "TLibEncoder/TEncEntropy.h"
#include "TLibEncoder/TEncSbac.h"
#define getStatistics getStatistics_reference( SAOStatData***, TComPicYuv*, TComPicYuv*, TComPic*, Bool ); Void getStatistics
#include "TLibEncoder/TEncSampleAdaptiveOffset.h"
#undef getStatistics
#endif
