#pragma once
#include "common.cuh"

#define BN_TABLE_MAX 24
struct BnDesc {
  const float *gamma, *beta, *rmean, *rvar;
  long long g_gs, b_gs;   // group strides of gamma/beta (sampled weights); running stats are shared
  int C, off;             // channels, offset inside the scale/shift scratch
};
struct BnTable { BnDesc d[BN_TABLE_MAX]; int n; };

int arpde_bn_fold_eval_launch(const BnTable& t, int groups, float* scale, float* shift, long long ss_gs, float eps, cudaStream_t stream);
int arpde_bn_finalize_train_launch(const double* stats, double count, const float* gamma, const float* beta, float* rmean,
                                   float* rvar, long long* nbt, float momentum, float eps, float* scale, float* shift,
                                   float* save_mean, float* save_invstd, int C, cudaStream_t stream);
