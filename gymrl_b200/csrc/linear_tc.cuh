// linear_tc.cuh — interface between linear.cu (dispatch) and linear_tc.cu (tcgen05 3xTF32 GEMM).
#pragma once
#include "common.cuh"

struct TcGemmParams {
    const float* A; int lda; const int32_t* a_rows;   // K-major: A[m][k] (rows gathered through a_rows); MN-major: A[k][m]
    const float* B; int ldb; const int32_t* b_rows;   // K-major: B[n][k];                               MN-major: B[k][n] (k rows gathered)
    float* C; int ldc;
    int M, N, K;
    const float* bias; int act;
    const float* H; int ldh; int act_in;
    int k_chunk;                 // reduction elements per blockIdx.z (multiple of 32)
    long long c_split_stride;
    float* colsum;               // MN-major A only (dW): colsum[blockIdx.z][M] = sum over this split's k of A[k][m] (nullable)
    int bn_max;                  // widest column tile to use (0 = 256): narrower tiles = fewer split-K partials for the dW GEMMs
};

bool tc_gemm_supported(const TcGemmParams& p, bool a_kmajor, bool b_kmajor);
int tc_gemm_launch(const TcGemmParams& p, bool a_kmajor, bool b_kmajor, int splits, cudaStream_t s);
