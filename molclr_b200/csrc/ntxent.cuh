// Internal interface between ntxent.cu (C-ABI entry points, striped backward) and ntxent_fused.cu (fused backward kernel).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace molclr {

constexpr int kNtxFusedMaxSplits = 32;      // partial-gradient slots the workspace reserves

int ntx_fused_splits(int64_t R, int64_t Rc);
size_t ntx_fused_ecol_floats(int64_t Rc);
// g partials [splits][R][C] of the NT-Xent backward for unit-norm fp16 rows with bounded logits (bound2 > 0), C <= 256
int ntx_bwd_fused(const __half* rep16, const __half* cols16, int ld16, int64_t R, int64_t Rc, int C, int64_t row_offset, int64_t row_offset2,
                  float inv_temperature, float bound2, const float* row_lse, const float* col_lse, float gscale, float* ecol, float* partials,
                  int splits, cudaStream_t stream);

}  // namespace molclr
