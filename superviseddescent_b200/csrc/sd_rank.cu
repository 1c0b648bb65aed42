// ColPivHouseholderQRSolver's diagnostic (regressors.hpp:288-293): numerical rank of the regularised A^T A.
//
// The reference runs Eigen::ColPivHouseholderQR on the D x D matrix only to ask rank() / isInvertible() and to print
// "The regularised AtA is not invertible ... (The rank is r, full rank would be D). Increase lambda."; the weights then come
// from the explicit inverse.  A^T A + Lambda is symmetric positive SEMI-definite, and for that class the rank-revealing
// factorisation is the diagonally pivoted Cholesky (LAPACK xPSTRF): at step k the largest remaining Schur-complement diagonal
// d_k is the pivot, and rank = #{k : d_k > threshold * d_0} with Eigen's default threshold eps * D.
//
// One CTA, left-looking, no row/column swaps (the pivot order is kept as an index list): column p of the current Schur complement
// is S[:,p] - sum_{m<k} L[:,m] L[p,m]; only the factor's rows (D x rank floats) are stored.  O(D^3 / 2) flops on one SM: a
// diagnostic path, like the reference's own ("much MUCH slower", regressors.hpp:242-243).  D <= kRankMaxDim.
#include "sd_internal.cuh"

#include <cstring>

namespace {

constexpr int kRankThreads = 1024;

__global__ void __launch_bounds__(kRankThreads) pivoted_cholesky_rank_kernel(const float* __restrict__ G, long long ldg, int D, float* __restrict__ Lt,
                                                                             float* __restrict__ dwork, float threshold, int* __restrict__ out /* rank, first pivot bits, last pivot bits */)
{
    __shared__ float red_v[kRankThreads / 32];
    __shared__ int red_i[kRankThreads / 32];
    __shared__ float s_piv;
    __shared__ int s_p;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // dwork[i] >= 0: remaining diagonal of column i; -1: already chosen
    for (int i = tid; i < D; i += kRankThreads) dwork[i] = G[(long long)i * ldg + i];
    __syncthreads();
    float d0 = 0.f, dlast = 0.f;
    int k = 0;
    for (; k < D; ++k) {
        float best = -1.f;
        int bi = -1;
        for (int i = tid; i < D; i += kRankThreads) {
            const float v = dwork[i];
            if (v > best) { best = v; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi >= 0 && (bi < 0 || oi < bi))) { best = ov; bi = oi; }
        }
        if (lane == 0) { red_v[warp] = best; red_i[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            best = lane < kRankThreads / 32 ? red_v[lane] : -1.f;
            bi = lane < kRankThreads / 32 ? red_i[lane] : -1;
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi >= 0 && (bi < 0 || oi < bi))) { best = ov; bi = oi; }
            }
            if (lane == 0) { s_piv = best; s_p = bi; }
        }
        __syncthreads();
        const float piv = s_piv;
        const int p = s_p;
        if (k == 0) d0 = piv;
        if (p < 0 || !(piv > threshold * d0) || !(piv > 0.f)) break;          // numerically zero from here on
        dlast = piv;
        const float r = sqrtf(piv);
        const float inv_r = 1.0f / r;
        float* Lk = Lt + (long long)k * D;
        // up to four rows per thread (D <= 4096): the m loop is shared, L[p, m] is loaded once per step for all of them
        float acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = tid + u * kRankThreads;
            acc[u] = 0.f;
            if (i < D) acc[u] = (i <= p) ? G[(long long)i * ldg + p] : G[(long long)p * ldg + i];   // symmetric: the upper triangle is stored
        }
        for (int m = 0; m < k; ++m) {
            const float* Lm = Lt + (long long)m * D;
            const float lp = Lm[p];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = tid + u * kRankThreads;
                if (i < D) acc[u] = fmaf(-Lm[i], lp, acc[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = tid + u * kRankThreads;
            if (i >= D) continue;
            const float di = dwork[i];
            float l = 0.f;
            if (i == p) l = r;
            else if (di >= 0.f) l = acc[u] * inv_r;
            Lk[i] = l;
            if (i == p) dwork[i] = -1.f;
            else if (di >= 0.f) { const float nd = di - l * l; dwork[i] = nd > 0.f ? nd : 0.f; }
        }
        __syncthreads();
    }
    if (tid == 0) { out[0] = k; out[1] = __float_as_int(d0); out[2] = __float_as_int(dlast); }
}

}  // namespace

// rank of the symmetric matrix whose upper triangle is in G (D x D, pitch ldg); -1 when D is beyond the diagnostic's range
int sd_gram_rank(sd_ctx* ctx, const float* d_G, int64_t ldg, int D, int* rank_out, float* first_pivot, float* last_pivot)
{
    *rank_out = -1;
    if (D > 4096) return SD_OK;
    float* ws = (float*)sd_workspace(ctx, SD_WS_PARTIAL, ((size_t)D * D + D) * sizeof(float) + 64);
    if (!ws) return SD_ERR_CUDA;
    int* d_out = reinterpret_cast<int*>(reinterpret_cast<char*>(ctx->d_scratch) + 128);
    const float threshold = 1.1920929e-7f * (float)D;                          // Eigen: NumTraits<float>::epsilon() * diagonalSize
    pivoted_cholesky_rank_kernel<<<1, kRankThreads, 0, ctx->stream>>>(d_G, ldg, D, ws + D, ws, threshold, d_out);
    SD_LAUNCH_CHECK(ctx, "pivoted_cholesky_rank_kernel");
    int* h = reinterpret_cast<int*>(reinterpret_cast<char*>(ctx->h_scratch) + 128);
    SD_CUDA(ctx, cudaMemcpyAsync(h, d_out, 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *rank_out = h[0];
    if (first_pivot) memcpy(first_pivot, &h[1], sizeof(float));
    if (last_pivot) memcpy(last_pivot, &h[2], sizeof(float));
    return SD_OK;
}
