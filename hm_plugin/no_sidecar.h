/* hm_plugin/no_sidecar.h -- force-included (after <cstdlib>) when compiling the reference's
 * App/TAppEncoder/encmain.cpp for the drop-in build: the two sidecar launches
 * (encmain.cpp:56 `system("python use_model.py")`, :106 `system("python gen_frames.py")`) become
 * no-ops, because the labels now come from libhevcdl.so inside TEncCu::compressCtu. */
#pragma once
#include <cstdlib>
#define system(cmd) ((void)(cmd), 0)
