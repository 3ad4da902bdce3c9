/* hm_plugin/rmd_hook.h -- force-included when compiling the reference's TEncSearch.cpp for the drop-in build.
 *
 * hm_plugin/Makefile wraps the two statements of the first-pass mode loop of TEncSearch::estIntraPredLumaQT that
 * compute one mode's SATD (HM_dl/source/Lib/TLibEncoder/TEncSearch.cpp:2303 predIntraAng(...) and :2306
 * uiSad += distParam.DistFunc(...)) in `if ( !hevcdl_hm_rmd_satd(...) ) { ... }` with sed at build time (the edited
 * copy lives in a temp dir and is never stored).  HEVCDL_RMD selects what the hook does:
 *   0  returns false: the reference code runs unchanged (byte-identical bitstreams);
 *   1  batched mode: supplies the SATD the B200 computed for that (PU, mode) against ORIGINAL-picture references for the
 *      whole frame up front (BD-rate clause); the two statements are skipped;
 *   2  exact mode: on the first mode of a PU hands the PU's original block and the RECONSTRUCTED reference samples HM just
 *      built (TComPattern.cpp:119-324, read back through TComPrediction::getPredictorPtr) to hevcdl_rmd_exact and serves the
 *      35 SATDs from that one call: bit-exact, so the bitstream is again byte-identical to the reference -- a parity
 *      demonstration of the device code inside the real encoder (one synchronous call per PU: not a speed-up).
 * Mode bits, lambda, xUpdateCandList and the MPM append stay the reference's in every mode. */
#pragma once
class TComDataCU;
class TComPrediction;
class TComTU;
/* Second hook (HEVCDL_TQ=1): the transform / quantisation / inverse-transform of one TU inside
 * TEncSearch::xIntraCodingTUBlock (TEncSearch.cpp:1301 m_pcTrQuant->transformNxN(...) through the inverse-transform
 * if/else ending before :1330 "//===== reconstruction =====") is wrapped in `if ( !hevcdl_hm_tu_code(...) ) { ... }` by a third
 * sed rule of hm_plugin/Makefile.  When the TU is one the device core covers -- flat quantiser, i.e. the encoder runs with
 * --RDOQ=0 --RDOQTS=0 --SignHideFlag=0, no transquant bypass -- or the rate-distortion optimised quantiser of the reference's
 * default options (RDOQ 1, RDOQTS 1, SignHideFlag 1: hevcdl_tu_code_rdoq, fed the CABAC bit-estimate table TEncSbac::estBit
 * has just filled, the component's lambda, the scan type and the cbf context) -- the hook sends the residual block to the
 * device (one synchronous call per TU: a parity demonstration like HEVCDL_RMD=2, not a speed-up), writes levels, cbf,
 * uiAbsSum and the reconstructed residual exactly where the reference code would, and the bitstream stays byte-identical to
 * the reference run with the same options.  Otherwise (RDOQ on with sign-bit hiding off or vice versa is covered too; transquant
 * bypass, scaling lists, non-square TUs are not) it returns false and the reference code runs. */
bool hevcdl_hm_tu_code( TComDataCU* pcCU, TComTU& rTu, int compID, short* piResi, unsigned uiStride, int* pcCoeff, int* puiAbsSum,
                        int qp, bool useTransformSkip, bool rdoqOn, const void* estBits );
bool hevcdl_hm_rmd_satd( TComPrediction* pred, TComDataCU* pcCU, unsigned x0InCu, unsigned y0InCu, unsigned width, unsigned mode,
                         const short* org, unsigned orgStride, unsigned* sad );
