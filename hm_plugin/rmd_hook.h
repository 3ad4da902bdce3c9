/* hm_plugin/rmd_hook.h -- force-included when compiling the reference's TEncSearch.cpp for the drop-in build.
 *
 * hm_plugin/Makefile wraps the two statements of the first-pass mode loop of TEncSearch::estIntraPredLumaQT that
 * compute one mode's SATD (HM_dl/source/Lib/TLibEncoder/TEncSearch.cpp:2303 predIntraAng(...) and :2306
 * uiSad += distParam.DistFunc(...)) in `if ( !hevcdl_hm_rmd_satd(...) ) { ... }` with sed at build time (the edited
 * copy lives in a temp dir and is never stored).  When the session runs with HEVCDL_RMD=1 the hook supplies the SATD
 * the B200 computed for that (PU, mode) against original-picture references and the two statements are skipped;
 * mode bits, lambda, xUpdateCandList and the MPM append stay the reference's.  With HEVCDL_RMD=0 it returns false
 * and the reference code runs unchanged. */
#pragma once
class TComDataCU;
bool hevcdl_hm_rmd_satd( TComDataCU* pcCU, unsigned x0InCu, unsigned y0InCu, unsigned width, unsigned mode, unsigned* sad );
