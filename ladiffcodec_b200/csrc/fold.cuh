// Load-time weight folds for the UNet (see fold.cu).
#pragma once
#include "common.cuh"

// out row stride out_ld (elements; 0 = Cin*K, i.e. dense rows)
int pack_conv_launch(const float* w, h16* out, int Cout, int Cin, int K, int standardize, cudaStream_t st, int out_ld = 0);
int pack_up_launch(const float* w, h16* out, int Cout, int Cin, cudaStream_t st);
int dup_bias_launch(const float* b, float* out, int n, cudaStream_t st);
int f32_to_bf16_launch(const float* x, h16* y, long long n, cudaStream_t st);
