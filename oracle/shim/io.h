/* oracle/shim/io.h -- TEST INFRASTRUCTURE (not product code).
 * POSIX stand-in for the MSVC <io.h> the reference encoder includes
 * (HM_dl/source/Lib/TLibEncoder/TEncCu.cpp:44,245 uses _access()). */
#pragma once
#include <unistd.h>
#define _access access
