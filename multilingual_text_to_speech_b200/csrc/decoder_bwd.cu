// Decoder backward (BPTT) for the teacher-forced decode: the reverse of decoder_fwd.cu.
// Restates what torch autograd replays for Decoder._decode (reference modules/tacotron2.py:148-209,
// train.py:83): frame/stop projection grads -> generator LSTM reverse loop -> attention LSTM +
// location-sensitive attention reverse loop (energies recomputed, never stored) -> time-batched dW GEMMs.
#include <cuda_bf16.h>
#include "decoder_internal.cuh"

namespace b200tts {

namespace {

inline int grid_for(size_t n) {
    size_t g = (n + 255) / 256;
    return (int)(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g));
}

// ---------------------------------------------------------------------------------------------
// utility kernels
// ---------------------------------------------------------------------------------------------
// dFS[i, b, n] = d_spec[b, i, n] (n < N), dFS[i, b, N] = d_stop[b, i]
__global__ void gather_frame_grads_kernel(float* __restrict__ dfs, const float* __restrict__ dspec,
                                          const float* __restrict__ dstop, int B, int T, int N) {
    const size_t total = (size_t)B * T * (N + 1);
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int n = idx % (N + 1);
        const int b = (idx / (N + 1)) % B;
        const int i = idx / ((size_t)(N + 1) * B);
        float v = 0.f;
        if (n < N) { if (dspec) v = dspec[((size_t)b * T + i) * N + n]; }
        else if (dstop) v = dstop[(size_t)b * T + i];
        dfs[idx] = v;
    }
}

// Column sums (bias gradients), deterministic two-level tree: partial[s][c] = sum of row slice s, then dst[c] += sum_s partial[s][c].
constexpr int COLSUM_SLICES = 64;
__global__ void colsum_partial_kernel(float* __restrict__ partial, const float* __restrict__ src, size_t rows, int cols, int ld) {
    __shared__ float sm[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const size_t per = (rows + gridDim.y - 1) / gridDim.y;
    const size_t r0 = blockIdx.y * per, r1 = r0 + per < rows ? r0 + per : rows;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (c < cols) {
        size_t r = r0 + threadIdx.y;
        for (; r + 24 < r1; r += 32) {
            a0 += src[r * ld + c]; a1 += src[(r + 8) * ld + c]; a2 += src[(r + 16) * ld + c]; a3 += src[(r + 24) * ld + c];
        }
        for (; r < r1; r += 8) a0 += src[r * ld + c];
    }
    sm[threadIdx.y][threadIdx.x] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += sm[j][threadIdx.x];
        partial[(size_t)blockIdx.y * cols + c] = s;
    }
}
__global__ void colsum_finish_kernel(float* __restrict__ dst, float* __restrict__ dst2, const float* __restrict__ partial, int slices, int cols) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float s = 0.f;
    for (int j = 0; j < slices; ++j) s += partial[(size_t)j * cols + c];
    dst[c] += s;
    if (dst2) dst2[c] += s;
}
// dst[c] (+ dst2[c]) += sum_r src[r*ld + c]; scratch holds COLSUM_SLICES * cols floats
int colsum_add(float* dst, float* dst2, const float* src, size_t rows, int cols, int ld, float* scratch, cudaStream_t st) {
    int slices = (int)(rows / 256);
    slices = slices < 1 ? 1 : (slices > COLSUM_SLICES ? COLSUM_SLICES : slices);
    colsum_partial_kernel<<<dim3(cdiv(cols, 32), slices), dim3(32, 8), 0, st>>>(scratch, src, rows, cols, ld);
    B200_LAUNCH_CHECK();
    colsum_finish_kernel<<<cdiv(cols, 128), 128, 0, st>>>(dst, dst2, scratch, slices, cols);
    B200_LAUNCH_CHECK();
    return B200TTS_OK;
}

// dst[j] += sum_b src[b*n + j]
__global__ void batchsum_add_kernel(float* __restrict__ dst, const float* __restrict__ src, int batch, size_t n) {
    for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int b = 0; b < batch; ++b) s += src[(size_t)b * n + j];
        dst[j] += s;
    }
}

// dst[r*ldd + c] += src[r*lds + c]
__global__ void add2d_kernel(float* __restrict__ dst, int ldd, const float* __restrict__ src, int lds, int rows, int cols) {
    const size_t total = (size_t)rows * cols;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int r = idx / cols, c = idx % cols;
        dst[(size_t)r * ldd + c] += src[(size_t)r * lds + c];
    }
}

// prenet layer backward through dropout + relu: dz = dy * scale * (y > 0)   (y is post relu+dropout)
__global__ void relu_dropout_bwd_kernel(float* __restrict__ dz, const float* __restrict__ dy, const float* __restrict__ y,
                                        float scale, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dz[i] = y[i] > 0.f ? dy[i] * scale : 0.f;
}

// ---------------------------------------------------------------------------------------------
// LSTM cell backward (pointwise) -- one thread owns 8 utterances of one hidden unit
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) lstm_cell_bwd_kernel(const CellBwdArgs p) {
    extern __shared__ __align__(16) float sm[];
    const int Bp = (p.B + 7) & ~7;
    float* dqT = sm;                                   // [A][Bp]
    float* wq = sm + (size_t)p.A * Bp;                 // [A][CELL_UNITS + 1]
    const int u0 = blockIdx.x * CELL_UNITS, D = p.D, A = p.A;
    if (p.dq) {
        for (int idx = threadIdx.x; idx < A * Bp; idx += blockDim.x) {
            const int b = idx / A, a = idx % A;
            dqT[a * Bp + b] = b < p.B ? p.dq[(size_t)b * A + a] : 0.f;
        }
        for (int idx = threadIdx.x; idx < A * CELL_UNITS; idx += blockDim.x) {
            const int a = idx / CELL_UNITS, uu = idx % CELL_UNITS;
            wq[a * (CELL_UNITS + 1) + uu] = (u0 + uu < D) ? p.Wq[(size_t)a * D + u0 + uu] : 0.f;
        }
        __syncthreads();
    }
    const float inv_h = 1.f / (1.f - p.rate_h), inv_c = 1.f / (1.f - p.rate_c);
    const int nbg = Bp / 8;
    for (int item = threadIdx.x; item < CELL_UNITS * nbg; item += blockDim.x) {
        const int uu = item % CELL_UNITS, bg = item / CELL_UNITS, u = u0 + uu;
        float dhq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) dhq[j] = 0.f;
        if (p.dq) {
            for (int a = 0; a < A; ++a) {
                const float w = wq[a * (CELL_UNITS + 1) + uu];
                const float4 d0 = *reinterpret_cast<const float4*>(&dqT[a * Bp + bg * 8]);
                const float4 d1 = *reinterpret_cast<const float4*>(&dqT[a * Bp + bg * 8 + 4]);
                dhq[0] = fmaf(w, d0.x, dhq[0]); dhq[1] = fmaf(w, d0.y, dhq[1]); dhq[2] = fmaf(w, d0.z, dhq[2]); dhq[3] = fmaf(w, d0.w, dhq[3]);
                dhq[4] = fmaf(w, d1.x, dhq[4]); dhq[5] = fmaf(w, d1.y, dhq[5]); dhq[6] = fmaf(w, d1.z, dhq[6]); dhq[7] = fmaf(w, d1.w, dhq[7]);
            }
        }
        if (u >= D) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int b = bg * 8 + j;
            if (b >= p.B) continue;
            const size_t bu = (size_t)b * D + u, g0 = (size_t)b * 4 * D + u;
            float dh = dhq[j];
            if (p.dh_static) dh += p.dh_static[(size_t)b * p.ld_dhs + u];
            float dc_in = 0.f, dh_rec = 0.f;
            if (!p.last) {
                for (int s = 0; s < p.nsplit; ++s) dh_rec += p.part[s * p.part_stride + (size_t)b * p.ld_part + p.part_col0 + u];
                dc_in = p.dc_state[bu];
                if (p.dhz_state) dh_rec += p.dhz_state[bu];
            }
            if (p.lengths && p.step >= p.lengths[b]) {    // frozen step: gradients of the state pass straight through
                p.dgates[g0] = 0.f; p.dgates[g0 + D] = 0.f; p.dgates[g0 + 2 * D] = 0.f; p.dgates[g0 + 3 * D] = 0.f;
                p.dc_state[bu] = dc_in;
                p.dhz_state[bu] = dh_rec;
                continue;
            }
            dh += dh_rec;
            const float gi = p.gates[g0], gf = p.gates[g0 + D], gg = p.gates[g0 + 2 * D], go = p.gates[g0 + 3 * D];
            const float cp = p.c_prev[bu];
            const float tc = tanhf(gf * cp + gi * gg);
            float dhn, dcn, dc_prev_direct = 0.f, dh_prev_direct = 0.f;     // grads wrt raw h', c'
            if (p.kind == B200TTS_CELL_ZONEOUT) {
                float kh, kc;    // d out / d raw
                if (p.training) {
                    kh = (1.f - p.rate_h) * (p.mask_h ? (float)p.mask_h[bu] * inv_h : 1.f);
                    kc = (1.f - p.rate_c) * (p.mask_c ? (float)p.mask_c[bu] * inv_c : 1.f);
                } else {
                    kh = 1.f - p.rate_h; kc = 1.f - p.rate_c;
                }
                dhn = dh * kh; dh_prev_direct = dh - dhn;
                dcn = dc_in * kc + dhn * go * (1.f - tc * tc);
                dc_prev_direct = dc_in - dc_in * kc;
            } else {
                dhn = (p.training && p.mask_h) ? dh * (float)p.mask_h[bu] * inv_h : dh;
                dcn = dc_in + dhn * go * (1.f - tc * tc);
            }
            p.dgates[g0] = dcn * gg * gi * (1.f - gi);
            p.dgates[g0 + D] = dcn * cp * gf * (1.f - gf);
            p.dgates[g0 + 2 * D] = dcn * gi * (1.f - gg * gg);
            p.dgates[g0 + 3 * D] = dhn * tc * go * (1.f - go);
            p.dc_state[bu] = dcn * gf + dc_prev_direct;
            if (p.dhz_state) p.dhz_state[bu] = dh_prev_direct;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Attention step backward, one CTA per utterance.  Recomputes location features and tanh
// arguments from the saved query and cumulative weights (SURVEY 7.3 "Backward memory").
// ---------------------------------------------------------------------------------------------
struct AttnBwdArgs {
    const float* q;            // [B, A] saved query of this step
    const float* memT;         // [B, L, A]
    const float* memory;       // [B, L, M]
    const int* lengths;
    const float* Wc; const float* Wloc; const float* bias; const float* v;
    const float* cum_prev;     // [B, L] cumulative weights the step consumed
    const float* w;  long long w_bstride;        // &align[0, i, 0]
    const float* dalign; long long dalign_bstride;   // &d_align[0, i, 0] or null
    const float* dctx_static;  // [B, M]
    const float* part; int nsplit; size_t part_stride; int ld_part;   // recurrent d ctx partials (cols [0, M)); null on last step
    float* dcum;               // [B, L] in: d cum_i, out: d cum_{i-1}
    float* dctx_tot;           // [B, M] out
    float* dq;                 // [B, A] out
    float* dmemT;              // [B, L, A] +=
    float* dWloc_acc;          // [B, A, C] +=
    float* dWc_acc;            // [B, C, K] +=
    float* dv_acc;             // [B, A] +=
    int B, L, M, A, C, K, LC, last;
};

struct AttnBwdSmem {
    int Lp, cumn, off_qb, off_vv, off_cump, off_Wl, off_WlT, off_Wcs, off_f, off_dF, off_w, off_de, off_ds, off_red, off_dctx,
        off_cred, total;
};
__host__ __device__ inline AttnBwdSmem attn_bwd_smem(int L, int M, int A, int C, int K, int LC) {
    AttnBwdSmem s;
    s.Lp = (L + 3) & ~3;
    s.cumn = (L + K - 1 + 3) & ~3;
    int o = 0;
    s.off_qb = o; o += A;
    s.off_vv = o; o += A;
    s.off_cump = o; o += s.cumn;
    s.off_Wl = o; o += A * C;
    s.off_WlT = o; o += C * A;
    s.off_Wcs = o; o += (C * K + 3) & ~3;
    s.off_f = o; o += C * s.Lp;
    s.off_dF = o; o += C * (s.Lp + 2 * K);        // K zeros either side so the transposed conv needs no bounds checks
    s.off_w = o; o += s.Lp;
    s.off_de = o; o += s.Lp;
    s.off_ds = o; o += LC * (A + 4);
    s.off_red = o; o += 64;
    s.off_dctx = o; o += (M + 3) & ~3;
    s.off_cred = o; o += (8 * A > 4 * s.Lp ? 8 * A : 4 * s.Lp);   // [8][A] query partials, later [4][Lp] conv partials
    s.total = o;
    return s;
}

__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_kernel(const AttnBwdArgs p) {
    extern __shared__ __align__(16) float sm[];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = ATT_THREADS / 32;
    const int L = p.L, A = p.A, C = p.C, K = p.K, M = p.M, LC = p.LC;
    const int half = (K - 1) / 2;
    const AttnBwdSmem so = attn_bwd_smem(L, M, A, C, K, LC);
    const int Lp = so.Lp, dFld = Lp + 2 * K, AS = A + 4;
    float* qb = sm + so.off_qb; float* vv = sm + so.off_vv; float* cump = sm + so.off_cump;
    float* Wl = sm + so.off_Wl; float* WlT = sm + so.off_WlT; float* Wcs = sm + so.off_Wcs;
    float* f = sm + so.off_f; float* dF = sm + so.off_dF; float* wS = sm + so.off_w; float* de = sm + so.off_de;
    float* dsS = sm + so.off_ds; float* red = sm + so.off_red; float* dctx = sm + so.off_dctx; float* cred = sm + so.off_cred;
    int len = p.lengths[b];
    len = len < 0 ? 0 : (len > L ? L : len);

    // ---- phase 1: stage small operands ----
    for (int a = tid; a < A; a += ATT_THREADS) { qb[a] = p.q[(size_t)b * A + a] + p.bias[a]; vv[a] = p.v[a]; }
    for (int j = tid; j < L + K - 1; j += ATT_THREADS) {
        const int l = j - half;
        cump[j] = (l >= 0 && l < L) ? p.cum_prev[(size_t)b * L + l] : 0.f;
    }
    for (int idx = tid; idx < A * C; idx += ATT_THREADS) {
        const float wv = p.Wloc[idx];
        Wl[idx] = wv;
        WlT[(idx % C) * A + idx / C] = wv;
    }
    for (int idx = tid; idx < C * K; idx += ATT_THREADS) Wcs[idx] = p.Wc[idx];
    for (int idx = tid; idx < C * dFld; idx += ATT_THREADS) dF[idx] = 0.f;
    for (int l = tid; l < Lp; l += ATT_THREADS) wS[l] = l < L ? p.w[(size_t)b * p.w_bstride + l] : 0.f;
    for (int m = tid; m < M; m += ATT_THREADS) {
        float g = p.dctx_static[(size_t)b * M + m];
        if (!p.last)
            for (int s = 0; s < p.nsplit; ++s) g += p.part[s * p.part_stride + (size_t)b * p.ld_part + m];
        dctx[m] = g;
        p.dctx_tot[(size_t)b * M + m] = g;
    }
    __syncthreads();

    // ---- phase 2: location features; dw[l] = dalign + dcum + <dctx, memory[l]> ----
    for (int idx = tid; idx < C * Lp; idx += ATT_THREADS) {
        const int c = idx / Lp, l = idx % Lp;
        float acc = 0.f;
        if (l < L)
            for (int k = 0; k < K; ++k) acc = fmaf(Wcs[c * K + k], cump[l + k], acc);
        f[idx] = acc;
    }
    for (int l = warp; l < Lp; l += NW) {
        float acc = 0.f;
        if (l < len) {
            const float* row = p.memory + ((size_t)b * L + l) * M;
            for (int m = lane; m < M; m += 32) acc = fmaf(dctx[m], row[m], acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            float g = 0.f;
            if (l < len) {
                g = acc + (p.last ? 0.f : p.dcum[(size_t)b * L + l]);
                if (p.dalign) g += p.dalign[(size_t)b * p.dalign_bstride + l];
            }
            de[l] = g;        // holds dw for now
        }
    }
    __syncthreads();

    // ---- phase 3: softmax backward ----
    float dot = 0.f;
    for (int l = tid; l < len; l += ATT_THREADS) dot = fmaf(wS[l], de[l], dot);
    dot = block_sum(dot, red);
    for (int l = tid; l < Lp; l += ATT_THREADS) de[l] = l < len ? wS[l] * (de[l] - dot) : 0.f;
    __syncthreads();

    // ---- phase 4: energy backward, chunked over positions ----
    float dq_reg[4] = {0.f, 0.f, 0.f, 0.f}, dv_reg[4] = {0.f, 0.f, 0.f, 0.f};
    const int ntile = ((A / 4) * (C / 4));                 // 4a x 4c tiles of dWloc (<= 256 by validation)
    const int t_a0 = (tid % (A / 4)) * 4, t_c0 = (tid / (A / 4)) * 4;
    float accW[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) accW[i][j] = 0.f;

    for (int lc0 = 0; lc0 < len; lc0 += LC) {
        const int rows = min(LC, ((len - lc0) + 3) & ~3);   // multiple of 4, rows beyond len are written as zeros
        // 4a: recompute s, ds
        for (int r0 = warp * 4; r0 < rows; r0 += NW * 4) {
            const int l0 = lc0 + r0;
            float s[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
            for (int c = 0; c < C; ++c) {
                const float4 fv = *reinterpret_cast<const float4*>(&f[c * Lp + l0]);
                float wv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) wv[j] = (lane + 32 * j < A) ? WlT[c * A + lane + 32 * j] : 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s[0][j] = fmaf(fv.x, wv[j], s[0][j]); s[1][j] = fmaf(fv.y, wv[j], s[1][j]);
                    s[2][j] = fmaf(fv.z, wv[j], s[2][j]); s[3][j] = fmaf(fv.w, wv[j], s[3][j]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int l = l0 + i;
                const float del = l < len ? de[l] : 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int a = lane + 32 * j;
                    if (a < A) {
                        float dsv = 0.f;
                        if (l < len) {
                            const size_t mi = ((size_t)b * L + l) * A + a;
                            const float th = tanhf(s[i][j] + qb[a] + p.memT[mi]);
                            dsv = del * vv[a] * (1.f - th * th);
                            dv_reg[j] = fmaf(del, th, dv_reg[j]);
                            dq_reg[j] += dsv;
                            p.dmemT[mi] += dsv;
                        }
                        dsS[(r0 + i) * AS + a] = dsv;
                    }
                }
            }
        }
        __syncthreads();
        // 4b: dF[c, l] = sum_a ds[l, a] * Wloc[a, c]      (4c x 4l register tiles)
        {
            const int ncg = C / 4;
            for (int t = tid; t < ncg * (rows / 4); t += ATT_THREADS) {
                const int c0 = (t % ncg) * 4, r0 = (t / ncg) * 4;
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
                for (int a = 0; a < A; ++a) {
                    const float4 w4 = *reinterpret_cast<const float4*>(&Wl[a * C + c0]);
                    const float d0 = dsS[(r0 + 0) * AS + a], d1 = dsS[(r0 + 1) * AS + a];
                    const float d2 = dsS[(r0 + 2) * AS + a], d3 = dsS[(r0 + 3) * AS + a];
                    acc[0][0] = fmaf(w4.x, d0, acc[0][0]); acc[0][1] = fmaf(w4.x, d1, acc[0][1]); acc[0][2] = fmaf(w4.x, d2, acc[0][2]); acc[0][3] = fmaf(w4.x, d3, acc[0][3]);
                    acc[1][0] = fmaf(w4.y, d0, acc[1][0]); acc[1][1] = fmaf(w4.y, d1, acc[1][1]); acc[1][2] = fmaf(w4.y, d2, acc[1][2]); acc[1][3] = fmaf(w4.y, d3, acc[1][3]);
                    acc[2][0] = fmaf(w4.z, d0, acc[2][0]); acc[2][1] = fmaf(w4.z, d1, acc[2][1]); acc[2][2] = fmaf(w4.z, d2, acc[2][2]); acc[2][3] = fmaf(w4.z, d3, acc[2][3]);
                    acc[3][0] = fmaf(w4.w, d0, acc[3][0]); acc[3][1] = fmaf(w4.w, d1, acc[3][1]); acc[3][2] = fmaf(w4.w, d2, acc[3][2]); acc[3][3] = fmaf(w4.w, d3, acc[3][3]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dF[(c0 + i) * dFld + K + lc0 + r0 + j] = acc[i][j];
            }
        }
        // 4c: dWloc[a, c] += sum_l ds[l, a] * f[c, l]       (4a x 4c register tile per thread, kept across chunks)
        if (tid < ntile) {
            for (int r = 0; r < rows; ++r) {
                const float4 d4 = *reinterpret_cast<const float4*>(&dsS[r * AS + t_a0]);
                const float f0 = f[(t_c0 + 0) * Lp + lc0 + r], f1 = f[(t_c0 + 1) * Lp + lc0 + r];
                const float f2 = f[(t_c0 + 2) * Lp + lc0 + r], f3 = f[(t_c0 + 3) * Lp + lc0 + r];
                accW[0][0] = fmaf(d4.x, f0, accW[0][0]); accW[0][1] = fmaf(d4.x, f1, accW[0][1]); accW[0][2] = fmaf(d4.x, f2, accW[0][2]); accW[0][3] = fmaf(d4.x, f3, accW[0][3]);
                accW[1][0] = fmaf(d4.y, f0, accW[1][0]); accW[1][1] = fmaf(d4.y, f1, accW[1][1]); accW[1][2] = fmaf(d4.y, f2, accW[1][2]); accW[1][3] = fmaf(d4.y, f3, accW[1][3]);
                accW[2][0] = fmaf(d4.z, f0, accW[2][0]); accW[2][1] = fmaf(d4.z, f1, accW[2][1]); accW[2][2] = fmaf(d4.z, f2, accW[2][2]); accW[2][3] = fmaf(d4.z, f3, accW[2][3]);
                accW[3][0] = fmaf(d4.w, f0, accW[3][0]); accW[3][1] = fmaf(d4.w, f1, accW[3][1]); accW[3][2] = fmaf(d4.w, f2, accW[3][2]); accW[3][3] = fmaf(d4.w, f3, accW[3][3]);
            }
        }
        __syncthreads();
    }
    if (tid < ntile) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) p.dWloc_acc[((size_t)b * A + t_a0 + i) * C + t_c0 + j] += accW[i][j];
    }
    // query / energy-vector gradients: reduce the per-warp partials
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int a = lane + 32 * j;
        if (a < A) { cred[warp * A + a] = dq_reg[j]; }
    }
    __syncthreads();
    for (int a = tid; a < A; a += ATT_THREADS) {
        float s = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < NW; ++w8) s += cred[w8 * A + a];
        p.dq[(size_t)b * A + a] = s;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int a = lane + 32 * j;
        if (a < A) { cred[warp * A + a] = dv_reg[j]; }
    }
    __syncthreads();
    for (int a = tid; a < A; a += ATT_THREADS) {
        float s = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < NW; ++w8) s += cred[w8 * A + a];
        p.dv_acc[(size_t)b * A + a] += s;
    }
    __syncthreads();

    // ---- phase 5: location conv backward ----
    // dWc[c, k] += sum_l dF[c, l] * cum[l + k - half]     (thread = one c, 4 consecutive k, sliding window)
    {
        const int nkg = (K + 3) / 4;
        for (int t = tid; t < C * nkg; t += ATT_THREADS) {
            const int c = t / nkg, k0 = (t % nkg) * 4;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            // cump index l + k; entries beyond the padded array are only reached for k >= K (discarded)
            float w0 = cump[min(k0, so.cumn - 1)], w1 = cump[min(k0 + 1, so.cumn - 1)], w2 = cump[min(k0 + 2, so.cumn - 1)];
            for (int l = 0; l < len; ++l) {
                const float w3 = cump[min(l + k0 + 3, so.cumn - 1)];
                const float d = dF[c * dFld + K + l];
                a0 = fmaf(d, w0, a0); a1 = fmaf(d, w1, a1); a2 = fmaf(d, w2, a2); a3 = fmaf(d, w3, a3);
                w0 = w1; w1 = w2; w2 = w3;
            }
            float* dst = p.dWc_acc + ((size_t)b * C + c) * K + k0;
            if (k0 < K) dst[0] += a0;
            if (k0 + 1 < K) dst[1] += a1;
            if (k0 + 2 < K) dst[2] += a2;
            if (k0 + 3 < K) dst[3] += a3;
        }
    }
    // d cum_{i-1}[j] = d cum_i[j] + sum_{c,k} dF[c, j + half - k] * Wc[c, k]   (thread = 4 consecutive j, a quarter of c)
    {
        const int njg = Lp / 4;
        for (int t = tid; t < njg * 4; t += ATT_THREADS) {
            const int jg = t % njg, cq = t / njg, j0 = jg * 4;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            for (int c = cq; c < C; c += 4) {
                const float* row = dF + c * dFld + K + half;     // row[j - k] = dF[c, j + half - k]; zero padding covers out-of-range
                float d1 = row[j0 + 1], d2 = row[j0 + 2], d3 = row[j0 + 3];
                for (int k = 0; k < K; ++k) {
                    const float d0 = row[j0 - k];
                    const float wv = Wcs[c * K + k];
                    a0 = fmaf(d0, wv, a0); a1 = fmaf(d1, wv, a1); a2 = fmaf(d2, wv, a2); a3 = fmaf(d3, wv, a3);
                    d3 = d2; d2 = d1; d1 = d0;
                }
            }
            cred[cq * Lp + j0] = a0; cred[cq * Lp + j0 + 1] = a1; cred[cq * Lp + j0 + 2] = a2; cred[cq * Lp + j0 + 3] = a3;
        }
    }
    __syncthreads();
    for (int j = tid; j < L; j += ATT_THREADS) {
        const float conv = cred[j] + cred[Lp + j] + cred[2 * Lp + j] + cred[3 * Lp + j];
        const float prev = p.last ? 0.f : p.dcum[(size_t)b * L + j];
        p.dcum[(size_t)b * L + j] = prev + conv;
    }
}

}  // namespace
int launch_cell_bwd(const CellBwdArgs& a, cudaStream_t st) {
    const int Bp = (a.B + 7) & ~7;
    const size_t smem = a.dq ? ((size_t)a.A * Bp + (size_t)a.A * (CELL_UNITS + 1)) * sizeof(float) : 0;
    static size_t configured = 48 * 1024;
    if (smem > configured) {
        B200_CUDA(cudaFuncSetAttribute(lstm_cell_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    lstm_cell_bwd_kernel<<<cdiv(a.D, CELL_UNITS), 256, smem, st>>>(a);
    B200_LAUNCH_CHECK();
    return B200TTS_OK;
}
namespace {

int pick_attn_bwd_chunk(int L, int M, int A, int C, int K) {
    const int Lp = (L + 3) & ~3;
    int LC = Lp;
    while (LC > 4 && (size_t)attn_bwd_smem(L, M, A, C, K, LC).total * sizeof(float) > 220 * 1024) LC -= 4;
    return LC;
}

int launch_attn_bwd(AttnBwdArgs a, cudaStream_t st) {
    a.LC = pick_attn_bwd_chunk(a.L, a.M, a.A, a.C, a.K);
    const size_t smem = (size_t)attn_bwd_smem(a.L, a.M, a.A, a.C, a.K, a.LC).total * sizeof(float);
    B200_REQUIRE(smem <= 227 * 1024, "attention backward: shared memory %zu B exceeds 227 KB (L=%d)", smem, a.L);
    static size_t configured = 48 * 1024;
    if (smem > configured) {
        B200_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    attn_bwd_kernel<<<a.B, ATT_THREADS, smem, st>>>(a);
    B200_LAUNCH_CHECK();
    return B200TTS_OK;
}

// ---------------------------------------------------------------------------------------------
// backward workspace
// ---------------------------------------------------------------------------------------------
struct BwdLayout {
    size_t dfs, dhgd, dctxs, dgg, dhas, dga, dq, dctxt, dcum, dc, dhz, dmemT, dWloc_acc, dWc_acc, dv_acc, dp1, dp0, dwfs,
        part, gpart, pextra, pextra2, dggb, dgab, total;      // dggb / dgab: bf16 [T, B, 4D] histories of the gate gradients (tcgen05 loops)
    int split_gen, split_att;
    size_t gpart_elems;
};

BwdLayout bwd_layout(const b200tts_decoder_shape& s) {
    BwdLayout l;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off = align_up(off + n, 64); return o; };
    const size_t T = s.T, B = s.B, D = s.D, M = s.M, P = s.P, A = s.A, N = s.N, L = s.L, C = s.C, K = s.K;
    l.dfs = take(T * B * (N + 1));
    l.dhgd = take(T * B * D);
    l.dctxs = take(T * B * M);
    l.dgg = take(T * B * 4 * D);
    l.dhas = take(T * B * D);
    l.dga = take(T * B * 4 * D);
    l.dq = take(T * B * A);
    l.dctxt = take(T * B * M);
    l.dcum = take(B * L);
    l.dc = take(B * D);
    l.dhz = take(B * D);
    l.dmemT = take(B * L * A);
    l.dWloc_acc = take(B * A * C);
    l.dWc_acc = take(B * C * K);
    l.dv_acc = take(B * A);
    l.dp1 = take(T * B * P);
    l.dp0 = take(T * B * P);
    l.dwfs = take((N + 1) * (D + M));
    l.split_gen = pick_splitk(s.B, s.D, 4 * s.D);
    l.split_att = pick_splitk(s.B, s.M + s.D, 4 * s.D);
    const size_t pg = (size_t)l.split_gen * B * D, pa = (size_t)l.split_att * B * (M + D);
    l.part = take(pg > pa ? pg : pa);
    // scratch for the split-K partials of the long-K weight-gradient GEMMs (only small outputs are split)
    l.gpart_elems = (size_t)6 * 1024 * 1024;
    l.gpart = take(l.gpart_elems);
    l.pextra = take(persist_bwd_gen_extra_bytes(s) / sizeof(float) + 64);
    l.pextra2 = take(att_bwd_extra(s).total / sizeof(float) + 64);
    l.dggb = take(T * B * 4 * D / 2 + 64);
    l.dgab = take(T * B * 4 * D / 2 + 64);
    l.total = off;
    return l;
}

// C (+)= op(A) . op(B), split-K chosen from the tile count; partial scratch shared by all calls
struct PackScope {
    PackScope() { tc_pack_cache_begin(); }
    ~PackScope() { tc_pack_cache_end(); }
};

int wgemm(cudaStream_t st, const BwdLayout& l, float* ws, int transA, int transB, int M, int N, int K, const float* A, int lda,
          const float* B, int ldb, float* C, int ldc, float beta, int batch = 1, long long sA = 0, long long sB = 0,
          long long sC = 0) {
    GemmDesc d;
    d.A = A; d.B = B; d.C = C; d.M = M; d.N = N; d.K = K; d.lda = lda; d.ldb = ldb; d.ldc = ldc; d.transA = transA;
    d.transB = transB; d.beta = beta; d.batch = batch; d.strideA = sA; d.strideB = sB; d.strideC = sC;
    return gemm_run_auto(d, ws + l.gpart, l.gpart_elems, st);
}
// weight gradient dW (+)= A^T . B with A [K, M] fp32 and B [K, N] fp32; B16 (optional) = the same B as bf16 rows (row stride ldb16) that
// the persistent forward loops left behind: read in place by the tcgen05 path (MN-major TMA operand), no conversion pass
int wgemm16(cudaStream_t st, const BwdLayout& l, float* ws, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
            const void* B16, int ldb16, float* C, int ldc, float beta, const void* A16 = nullptr, int lda16 = 0) {
    GemmDesc d;
    d.A = A; d.B = B; d.C = C; d.M = M; d.N = N; d.K = K; d.lda = lda; d.ldb = ldb; d.ldc = ldc; d.transA = 1; d.transB = 0; d.beta = beta;
    d.B16 = B16; d.ldb16 = ldb16; d.A16 = A16; d.lda16 = lda16;
    return gemm_run_auto(d, ws + l.gpart, l.gpart_elems, st);
}
// input gradient dX = A . B with A [M, K] fp32 (A16: the same matrix as bf16 rows, read in place) and B [K, N] fp32
int xgemm16(cudaStream_t st, const BwdLayout& l, float* ws, int M, int N, int K, const float* A, int lda, const void* A16, int lda16,
            const float* B, int ldb, float* C, int ldc, float beta) {
    GemmDesc d;
    d.A = A; d.B = B; d.C = C; d.M = M; d.N = N; d.K = K; d.lda = lda; d.ldb = ldb; d.ldc = ldc; d.transA = 0; d.transB = 0; d.beta = beta;
    d.A16 = A16; d.lda16 = lda16;
    return gemm_run_auto(d, ws + l.gpart, l.gpart_elems, st);
}

}  // namespace

size_t decoder_bwd_workspace_floats(const b200tts_decoder_shape& s) { return bwd_layout(s).total; }
// byte offset of the phase counters of the persistent backward kernels inside the backward workspace (0: generator, 1: attention)
size_t decoder_bwd_profile_offset(const b200tts_decoder_shape& s, int which) {
    const BwdLayout l = bwd_layout(s);
    if (which == 0) return l.pextra * sizeof(float) + persist_bwd_gen_extra_bytes(s) - 148 * 8 * 8;
    return l.pextra2 * sizeof(float) + att_bwd_extra(s).barrier + 256;
}

// standalone backward of one attention step (module-level API): per-utterance accumulators in the workspace, reduced over the batch here
size_t attention_step_backward_workspace_elems(int B, int M, int A, int C, int K) {
    return (size_t)B * ((size_t)A * C + (size_t)C * K + A + M);
}
int attention_step_backward_impl(int B, int L, int M, int A, int C, int K, const float* q, const float* memory, const float* memT,
                                 const int* lengths, const float* Wloc, const float* Wc, const float* bias, const float* v,
                                 const float* cum_prev, const float* weights, const float* d_ctx, const float* d_weights, float* d_cum,
                                 float* d_q, float* d_memT, float* d_Wloc, float* d_Wc, float* d_v, float* ws, cudaStream_t st) {
    B200_REQUIRE(A <= 128 && A % 4 == 0 && C % 4 == 0 && (A / 4) * (C / 4) <= ATT_THREADS && M <= 512 && (K % 2) == 1,
                 "attention_step_backward: unsupported dims A=%d C=%d M=%d K=%d", A, C, M, K);
    float* dWloc_acc = ws;
    float* dWc_acc = dWloc_acc + (size_t)B * A * C;
    float* dv_acc = dWc_acc + (size_t)B * C * K;
    float* dctx_tot = dv_acc + (size_t)B * A;
    B200_TRY(launch_fill(ws, 0.f, (size_t)B * ((size_t)A * C + (size_t)C * K + A), st));
    AttnBwdArgs aa{};
    aa.q = q; aa.memT = memT; aa.memory = memory; aa.lengths = lengths;
    aa.Wc = Wc; aa.Wloc = Wloc; aa.bias = bias; aa.v = v; aa.cum_prev = cum_prev;
    aa.w = weights; aa.w_bstride = L; aa.dalign = d_weights; aa.dalign_bstride = L;
    aa.dctx_static = d_ctx; aa.part = nullptr; aa.nsplit = 0; aa.part_stride = 0; aa.ld_part = 0;
    aa.dcum = d_cum; aa.dctx_tot = dctx_tot; aa.dq = d_q; aa.dmemT = d_memT;
    aa.dWloc_acc = dWloc_acc; aa.dWc_acc = dWc_acc; aa.dv_acc = dv_acc;
    aa.B = B; aa.L = L; aa.M = M; aa.A = A; aa.C = C; aa.K = K; aa.last = 0;
    B200_TRY(launch_attn_bwd(aa, st));
    batchsum_add_kernel<<<grid_for((size_t)A * C), 256, 0, st>>>(d_Wloc, dWloc_acc, B, (size_t)A * C);
    B200_LAUNCH_CHECK();
    batchsum_add_kernel<<<grid_for((size_t)C * K), 256, 0, st>>>(d_Wc, dWc_acc, B, (size_t)C * K);
    B200_LAUNCH_CHECK();
    batchsum_add_kernel<<<1, 256, 0, st>>>(d_v, dv_acc, B, (size_t)A);
    B200_LAUNCH_CHECK();
    return B200TTS_OK;
}

int decoder_backward_impl(const b200tts_decoder_shape& s, const b200tts_decoder_params& w, const b200tts_decoder_inputs& in,
                          const b200tts_decoder_outputs& fwd_out, const b200tts_decoder_output_grads& dout, const float* fws,
                          float* bws, size_t bws_bytes, const b200tts_decoder_params& dw, float* d_memory, cudaStream_t st) {
    B200_TRY(validate_decoder_shape(s));
    B200_REQUIRE(fwd_out.alignments, "decoder_backward: the forward alignments tensor is required");
    B200_REQUIRE(s.A % 4 == 0 && s.C % 4 == 0 && (s.A / 4) * (s.C / 4) <= ATT_THREADS,
                 "decoder_backward: attention dims A=%d C=%d unsupported (need A%%4==0, C%%4==0, A*C<=4096)", s.A, s.C);
    if (in.teacher)
        for (int i = 0; i < s.T; ++i)
            if (!in.teacher[i]) {
                set_last_error("decoder_backward: step %d is free-running; backward through free-running steps is not implemented", i);
                return B200TTS_ERR_UNSUPPORTED;
            }
    const DecoderLayout fl = decoder_layout(s);
    const BwdLayout l = bwd_layout(s);
    B200_REQUIRE(bws_bytes >= l.total * sizeof(float), "decoder_backward: workspace too small (%zu < %zu bytes)", bws_bytes,
                 l.total * sizeof(float));
    const int B = s.B, T = s.T, D = s.D, M = s.M, P = s.P, N = s.N, A = s.A, L = s.L, C = s.C, K = s.K, MD = M + D, N1 = N + 1;
    const size_t BD = (size_t)B * D, B4D = 4 * BD, TB = (size_t)T * B;
    auto F = [&](size_t off) { return fws + off; };
    auto W = [&](size_t off) { return bws + off; };
    const float* ai = F(fl.ai);                // [T+1, B, M+D]
    const float* ai1 = ai + (size_t)B * MD;    // rows 1..T
    // bf16 operand rows the tcgen05 forward loops left in the persistent workspace: aib [T+1, B, Kp_att] = [h_att | ctx | 0], hgb [T+1, B, Kp_gen]
    // = h_gen (row i+1 = state after step i, row 0 = 0).  The weight-gradient products read them in place (MN-major TMA operands).
    const bool tc_rows = precision_mode() == B200TTS_PRECISION_BF16 && s.training && tc_persist_supported(s) && persist_att_bwd_supported(s);
    const PersistLayout prl = persist_layout(s);
    const TcPersistGeom tcg = tc_persist_geom(s);
    const unsigned char* pws_rows = reinterpret_cast<const unsigned char*>(F(fl.persist));
    const __nv_bfloat16* aib = tc_rows ? reinterpret_cast<const __nv_bfloat16*>(pws_rows + prl.aib) : nullptr;
    const __nv_bfloat16* hgb = tc_rows ? reinterpret_cast<const __nv_bfloat16*>(pws_rows + prl.hgb) : nullptr;
    const int ldab = tcg.Kp_att, ldhb = tcg.Kp_gen;
    const __nv_bfloat16* aib1 = aib ? aib + (size_t)B * ldab : nullptr;      // rows 1..T
    const __nv_bfloat16* hgb1 = hgb ? hgb + (size_t)B * ldhb : nullptr;

    // ---- 1. frame / stop projection backward (time-batched) ----
    gather_frame_grads_kernel<<<grid_for(TB * N1), 256, 0, st>>>(W(l.dfs), dout.d_spectrogram, dout.d_stop, B, T, N);
    B200_LAUNCH_CHECK();
    // d h_gen (direct) and d ctx (projection part)
    B200_TRY(wgemm(st, l, bws, 0, 0, (int)TB, D, N1, W(l.dfs), N1, F(fl.wfs), D + M, W(l.dhgd), D, 0.f));
    B200_TRY(wgemm(st, l, bws, 0, 0, (int)TB, M, N1, W(l.dfs), N1, F(fl.wfs) + D, D + M, W(l.dctxs), M, 0.f));
    // d [frame_w ; stop_w] = dFS^T . [h_gen | ctx]
    B200_TRY(wgemm16(st, l, bws, N1, D, (int)TB, W(l.dfs), N1, F(fl.hg) + BD, D, hgb1, ldhb, W(l.dwfs), D + M, 0.f));
    B200_TRY(wgemm16(st, l, bws, N1, M, (int)TB, W(l.dfs), N1, ai1, MD, aib1 ? aib1 + D : nullptr, ldab, W(l.dwfs) + D, D + M, 0.f));
    add2d_kernel<<<grid_for((size_t)N * (D + M)), 256, 0, st>>>(dw.frame_w, D + M, W(l.dwfs), D + M, N, D + M);
    B200_LAUNCH_CHECK();
    add2d_kernel<<<grid_for((size_t)(D + M)), 256, 0, st>>>(dw.stop_w, D + M, W(l.dwfs) + (size_t)N * (D + M), D + M, 1, D + M);
    B200_LAUNCH_CHECK();
    B200_TRY(colsum_add(dw.frame_b, nullptr, W(l.dfs), TB, N, N1, W(l.gpart), st));
    B200_TRY(colsum_add(dw.stop_b, nullptr, W(l.dfs) + N, TB, 1, N1, W(l.gpart), st));

    // ---- 2. generator LSTM reverse loop ----
    const bool zone = s.cell_kind == B200TTS_CELL_ZONEOUT;
    // the tcgen05 reverse loops keep their bf16 gate gradients as [T, B, 4D] histories: the time-batched products below read them in place
    // (K-major for dX, MN-major for dW) instead of converting the fp32 copies
    const bool hist_gen = precision_mode() == B200TTS_PRECISION_BF16 && persist_bwd_supported(s) && tc_persist_gen_bwd_supported(s) &&
                          !getenv("B200TTS_NO_DGB_HISTORY");
    void* dggb = hist_gen ? static_cast<void*>(W(l.dggb)) : nullptr;
    if (precision_mode() == B200TTS_PRECISION_BF16 && persist_bwd_supported(s)) {
        // bf16 perf mode: one cooperative weight-stationary kernel for the whole reverse recurrence
        if (tc_persist_gen_bwd_supported(s))      // TMA + tcgen05 + TMEM variant (decoder_persist_bwd_tc.cu)
            B200_TRY(tc_persist_gen_bwd_loop(s, w, in, fl, fws, W(l.dhgd), W(l.dgg), reinterpret_cast<unsigned char*>(W(l.pextra)), st, dggb));
        else
            B200_TRY(persist_gen_bwd_loop(s, w, in, fl, fws, W(l.dhgd), W(l.dgg), reinterpret_cast<unsigned char*>(W(l.pextra)), st));
    } else {
    for (int i = T - 1; i >= 0; --i) {
            CellBwdArgs ca{};
            ca.gates = F(fl.gg) + (size_t)i * B4D;
            ca.c_prev = F(fl.cg) + (size_t)i * BD;
            ca.dh_static = W(l.dhgd) + (size_t)i * BD; ca.ld_dhs = D;
            ca.part = W(l.part); ca.nsplit = l.split_gen; ca.part_stride = BD; ca.ld_part = D; ca.part_col0 = 0;
            ca.dq = nullptr; ca.Wq = nullptr; ca.A = 0;
            ca.dc_state = W(l.dc); ca.dhz_state = zone ? W(l.dhz) : nullptr;
            ca.mask_h = in.mask_gen_h ? in.mask_gen_h + (size_t)i * BD : nullptr;
            ca.mask_c = in.mask_gen_c ? in.mask_gen_c + (size_t)i * BD : nullptr;
            ca.kind = s.cell_kind; ca.training = s.training; ca.rate_h = s.rate_h; ca.rate_c = s.rate_c;
            ca.dgates = W(l.dgg) + (size_t)i * B4D; ca.B = B; ca.D = D; ca.last = (i == T - 1);
            B200_TRY(launch_cell_bwd(ca, st));
            if (i > 0) {
                GemmDesc d;      // d h_gen_{i-1} (recurrent) = dgates_i . W_hh
                d.A = ca.dgates; d.lda = 4 * D; d.B = w.gen_w_hh; d.ldb = D; d.transB = 0; d.M = B; d.N = D; d.K = 4 * D;
                d.splitk = l.split_gen; d.partial = W(l.part); d.keep_partials = 1;
                if (d.splitk == 1) { d.C = W(l.part); d.ldc = D; d.keep_partials = 0; d.partial = nullptr; }
                B200_TRY(gemm_run(d, st));
            }
        }
    }
    {
        // time-batched generator gradients.  The gate gradients are final now: their packed (transposed / K-contiguous) bf16 copies are
        // made once and shared by the three weight-gradient and the two input-gradient products (pack cache of the tcgen05 GEMM).
        PackScope pack_scope;
        B200_TRY(wgemm16(st, l, bws, 4 * D, D, (int)TB, W(l.dgg), 4 * D, F(fl.hg), D, hgb, ldhb, dw.gen_w_hh, D, 1.f, dggb, 4 * D));
        B200_TRY(wgemm16(st, l, bws, 4 * D, D, (int)TB, W(l.dgg), 4 * D, ai1 + M, MD, aib1, ldab, dw.gen_w_ih, D + M, 1.f, dggb, 4 * D));
        B200_TRY(wgemm16(st, l, bws, 4 * D, M, (int)TB, W(l.dgg), 4 * D, ai1, MD, aib1 ? aib1 + D : nullptr, ldab, dw.gen_w_ih + D, D + M, 1.f, dggb, 4 * D));
        B200_TRY(colsum_add(dw.gen_b_ih, dw.gen_b_hh, W(l.dgg), TB, 4 * D, 4 * D, W(l.gpart), st));
        // d h_att (static part) and d ctx (generator-input part, accumulated onto the projection part)
        B200_TRY(xgemm16(st, l, bws, (int)TB, D, 4 * D, W(l.dgg), 4 * D, dggb, 4 * D, w.gen_w_ih, D + M, W(l.dhas), D, 0.f));
        B200_TRY(xgemm16(st, l, bws, (int)TB, M, 4 * D, W(l.dgg), 4 * D, dggb, 4 * D, w.gen_w_ih + D, D + M, W(l.dctxs), M, 1.f));
    }

    // ---- 3. attention LSTM + attention reverse loop ----
    const bool persist_att = precision_mode() == B200TTS_PRECISION_BF16 && s.training && (tc_persist_supported(s) || persist_supported(s)) &&
                             persist_att_bwd_supported(s);
    void* dgab = (persist_att && persist_att_bwd_tc(s) && !getenv("B200TTS_NO_DGB_HISTORY")) ? static_cast<void*>(W(l.dgab)) : nullptr;
    if (persist_att) {
        // bf16 perf mode: cooperative weight-stationary kernel (tensor-core attention backward inside), then a parallel post pass
        const PersistLayout pl = persist_layout(s);
        B200_TRY(persist_att_bwd_loop(s, w, in, fl, fws, pl, reinterpret_cast<const unsigned char*>(F(fl.persist)), fwd_out.alignments,
                                      dout.d_alignments, W(l.dhas), W(l.dctxs), W(l.dga), W(l.dq), W(l.dctxt), W(l.dmemT),
                                      reinterpret_cast<unsigned char*>(W(l.pextra2)), dw, st, dgab));
    } else {
    B200_TRY(launch_fill(W(l.dmemT), 0.f, (size_t)B * L * A, st));
        B200_TRY(launch_fill(W(l.dWloc_acc), 0.f, (size_t)B * A * C, st));
        B200_TRY(launch_fill(W(l.dWc_acc), 0.f, (size_t)B * C * K, st));
        B200_TRY(launch_fill(W(l.dv_acc), 0.f, (size_t)B * A, st));
        for (int i = T - 1; i >= 0; --i) {
            const int last = (i == T - 1);
            AttnBwdArgs aa{};
            aa.q = F(fl.q) + (size_t)i * B * A; aa.memT = F(fl.memT); aa.memory = in.memory; aa.lengths = in.text_lengths;
            aa.Wc = w.attn_loc_features; aa.Wloc = w.attn_location; aa.bias = w.attn_bias; aa.v = w.attn_energy;
            aa.cum_prev = F(fl.cum) + (size_t)i * B * L;
            aa.w = fwd_out.alignments + (size_t)i * L; aa.w_bstride = (long long)T * L;
            aa.dalign = dout.d_alignments ? dout.d_alignments + (size_t)i * L : nullptr; aa.dalign_bstride = (long long)T * L;
            aa.dctx_static = W(l.dctxs) + (size_t)i * B * M;
            aa.part = W(l.part); aa.nsplit = l.split_att; aa.part_stride = (size_t)B * MD; aa.ld_part = MD;
            aa.dcum = W(l.dcum); aa.dctx_tot = W(l.dctxt) + (size_t)i * B * M; aa.dq = W(l.dq) + (size_t)i * B * A;
            aa.dmemT = W(l.dmemT); aa.dWloc_acc = W(l.dWloc_acc); aa.dWc_acc = W(l.dWc_acc); aa.dv_acc = W(l.dv_acc);
            aa.B = B; aa.L = L; aa.M = M; aa.A = A; aa.C = C; aa.K = K; aa.last = last;
            B200_TRY(launch_attn_bwd(aa, st));

            CellBwdArgs ca{};
            ca.gates = F(fl.ga) + (size_t)i * B4D;
            ca.c_prev = F(fl.ca) + (size_t)i * BD;
            ca.dh_static = W(l.dhas) + (size_t)i * BD; ca.ld_dhs = D;
            ca.part = W(l.part); ca.nsplit = l.split_att; ca.part_stride = (size_t)B * MD; ca.ld_part = MD; ca.part_col0 = M;
            ca.dq = aa.dq; ca.Wq = w.attn_query; ca.A = A;
            ca.dc_state = W(l.dc); ca.dhz_state = zone ? W(l.dhz) : nullptr;
            ca.mask_h = in.mask_att_h ? in.mask_att_h + (size_t)i * BD : nullptr;
            ca.mask_c = in.mask_att_c ? in.mask_att_c + (size_t)i * BD : nullptr;
            ca.kind = s.cell_kind; ca.training = s.training; ca.rate_h = s.rate_h; ca.rate_c = s.rate_c;
            ca.dgates = W(l.dga) + (size_t)i * B4D; ca.B = B; ca.D = D; ca.last = last;
            B200_TRY(launch_cell_bwd(ca, st));
            if (i > 0) {
                GemmDesc d;      // [d ctx_{i-1} | d h_att_{i-1}] (recurrent) = dgates_i . [W_ih[:, P:] | W_hh]
                d.A = ca.dgates; d.lda = 4 * D; d.B = F(fl.wcat_att); d.ldb = MD; d.transB = 0; d.M = B; d.N = MD; d.K = 4 * D;
                d.splitk = l.split_att; d.partial = W(l.part); d.keep_partials = 1;
                if (d.splitk == 1) { d.C = W(l.part); d.ldc = MD; d.keep_partials = 0; d.partial = nullptr; }
                B200_TRY(gemm_run(d, st));
            }
        }
    }

    // ---- 4. time-batched gradients of the attention LSTM, attention parameters, prenet, memory ----
    {
        PackScope pack_scope;       // one transposed bf16 copy of the attention-LSTM gate gradients for the three weight-gradient products
        B200_TRY(wgemm16(st, l, bws, 4 * D, P, (int)TB, W(l.dga), 4 * D, F(fl.p1), P, nullptr, 0, dw.att_w_ih, P + M, 1.f, dgab, 4 * D));
        B200_TRY(wgemm16(st, l, bws, 4 * D, M, (int)TB, W(l.dga), 4 * D, ai, MD, aib ? aib + D : nullptr, ldab, dw.att_w_ih + P, P + M, 1.f, dgab, 4 * D));
        B200_TRY(wgemm16(st, l, bws, 4 * D, D, (int)TB, W(l.dga), 4 * D, ai + M, MD, aib, ldab, dw.att_w_hh, D, 1.f, dgab, 4 * D));
    }
    {
        B200_TRY(colsum_add(dw.att_b_ih, dw.att_b_hh, W(l.dga), TB, 4 * D, 4 * D, W(l.gpart), st));
        B200_TRY(colsum_add(dw.attn_bias, nullptr, W(l.dq), TB, A, A, W(l.gpart), st));
    }
    // d Wq = dQ^T . h_att
    B200_TRY(wgemm16(st, l, bws, A, D, (int)TB, W(l.dq), A, ai1 + M, MD, aib1, ldab, dw.attn_query, D, 1.f));
    if (!persist_att) {
        batchsum_add_kernel<<<grid_for((size_t)A * C), 256, 0, st>>>(dw.attn_location, W(l.dWloc_acc), B, (size_t)A * C);
        B200_LAUNCH_CHECK();
        batchsum_add_kernel<<<grid_for((size_t)C * K), 256, 0, st>>>(dw.attn_loc_features, W(l.dWc_acc), B, (size_t)C * K);
        B200_LAUNCH_CHECK();
        batchsum_add_kernel<<<1, 256, 0, st>>>(dw.attn_energy, W(l.dv_acc), B, (size_t)A);
        B200_LAUNCH_CHECK();
    }
    // d Wm = dmemT^T . memory ; d memory = align^T . dctx (per utterance) + dmemT . Wm
    B200_TRY(wgemm(st, l, bws, 1, 0, A, M, B * L, W(l.dmemT), A, in.memory, M, dw.attn_memory, M, 1.f));
    if (d_memory) {
        B200_TRY(wgemm(st, l, bws, 1, 0, L, M, T, fwd_out.alignments, L, W(l.dctxt), B * M, d_memory, M, 0.f, B,
                       (long long)T * L, (long long)M, (long long)L * M));
        B200_TRY(wgemm(st, l, bws, 0, 0, B * L, M, A, W(l.dmemT), A, w.attn_memory, M, d_memory, M, 1.f));
    }
    // prenet: d P1 = dGA . W_ih[:, :P]; through dropout+relu; layer 1; layer 0
    {
        const float scale1 = in.mask_prenet1 ? 1.f / (1.f - s.prenet_rate) : 1.f;
        const float scale0 = in.mask_prenet0 ? 1.f / (1.f - s.prenet_rate) : 1.f;
        B200_TRY(xgemm16(st, l, bws, (int)TB, P, 4 * D, W(l.dga), 4 * D, dgab, 4 * D, w.att_w_ih, P + M, W(l.dp1), P, 0.f));
        relu_dropout_bwd_kernel<<<grid_for(TB * P), 256, 0, st>>>(W(l.dp1), W(l.dp1), F(fl.p1), scale1, TB * P);
        B200_LAUNCH_CHECK();
        B200_TRY(wgemm(st, l, bws, 1, 0, P, P, (int)TB, W(l.dp1), P, F(fl.p0), P, dw.prenet_w1, P, 1.f));
        B200_TRY(colsum_add(dw.prenet_b1, nullptr, W(l.dp1), TB, P, P, W(l.gpart), st));
        B200_TRY(wgemm(st, l, bws, 0, 0, (int)TB, P, P, W(l.dp1), P, w.prenet_w1, P, W(l.dp0), P, 0.f));
        relu_dropout_bwd_kernel<<<grid_for(TB * P), 256, 0, st>>>(W(l.dp0), W(l.dp0), F(fl.p0), scale0, TB * P);
        B200_LAUNCH_CHECK();
        B200_TRY(wgemm(st, l, bws, 1, 0, P, N, (int)TB, W(l.dp0), P, F(fl.xtm), N, dw.prenet_w0, N, 1.f));
        B200_TRY(colsum_add(dw.prenet_b0, nullptr, W(l.dp0), TB, P, P, W(l.gpart), st));
    }
    return B200TTS_OK;
}

}  // namespace b200tts
