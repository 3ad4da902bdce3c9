// hm_plugin/ref_pred_call.cpp -- compiled with -DpredIntraAng=predIntraAng_reference (hm_plugin/Makefile): inside this
// translation unit the class declares the renamed member, so this is a call of the reference's own intra predictor
// (HM_dl/source/Lib/TLibCommon/TComPrediction.cpp:390-472), used by TComPrediction_hevcdl.cpp when HEVCDL_PRED is off or the
// block is not one the device predictor covers.
#include "TLibCommon/TComPrediction.h"
#include "TLibCommon/TComTU.h"

void hevcdl_ref_predIntraAng( TComPrediction *p, const ComponentID compID, UInt uiDirMode, Pel *piOrg, UInt uiOrgStride, Pel *piPred, UInt uiStride,
                              TComTU &rTu, const Bool bUseFilteredPredSamples, const Bool bUseLosslessDPCM )
{
  p->predIntraAng( compID, uiDirMode, piOrg, uiOrgStride, piPred, uiStride, rTu, bUseFilteredPredSamples, bUseLosslessDPCM );
}
