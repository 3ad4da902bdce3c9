/* oracle/shim/Windows.h -- TEST INFRASTRUCTURE (not product code).
 * POSIX stand-in for <Windows.h>: the reference only needs Sleep(ms)
 * (HM_dl/source/Lib/TLibEncoder/TEncCu.cpp:45,245). */
#pragma once
#include <unistd.h>
static inline void Sleep(unsigned ms) { usleep(ms * 1000u); }
