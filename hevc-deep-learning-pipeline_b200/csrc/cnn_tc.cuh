// cnn_tc.cuh -- tcgen05 (5th-gen tensor core) implementation of the depth-prediction CNN.
// PLACEHOLDER until the tensor-core kernels land: HEVCDL_PREC_BF16_TC is rejected at create time.
#pragma once
#include <string>

#include "common.cuh"

namespace hevcdl {
struct TcParams { int ready = 0; };
inline int tc_prepare_weights(const float *, const float *, TcParams *, void **, std::string &) { return HEVCDL_OK; }
inline int tc_configure(std::string &) { return HEVCDL_OK; }
constexpr bool TC_BUILT = false;
inline int tc_launch(const TcParams &, const uint8_t *, const uint8_t *, const uint8_t *, FrameGeom, int, int, int,
                     uint8_t *, float *, int, cudaStream_t) { return 0; }
}  // namespace hevcdl
