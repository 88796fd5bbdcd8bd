// comm.cuh — internal NCCL helpers (see comm.cu).
#pragma once
#include "common.cuh"
namespace mvs {
int comm_allreduce_sum_f32(mvs_ctx* ctx, float* buf_dev, size_t count);
int comm_allreduce_sum_f64(mvs_ctx* ctx, double* buf_dev, size_t count);
int comm_allreduce_sum_i64(mvs_ctx* ctx, long long* buf_dev, size_t count);
}  // namespace mvs
