// hm_plugin/ref_loopfilter_call.cpp -- compiled with -DloopFilterPic=loopFilterPic_reference (hm_plugin/Makefile): inside this
// translation unit the class declares the renamed member, so this is a call of the reference's own deblocking filter body
// (HM_dl/source/Lib/TLibCommon/TComLoopFilter.cpp:130-158), used by TComLoopFilter_hevcdl.cpp when HEVCDL_DBF is off or the
// picture is not one the device filter covers.
#include "TLibCommon/TComLoopFilter.h"
#include "TLibCommon/TComPic.h"

void hevcdl_ref_loopFilterPic( TComLoopFilter *lf, TComPic *pcPic ) { lf->loopFilterPic( pcPic ); }
