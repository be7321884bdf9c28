// Persistent recurrent kernels of the bf16 perf mode (backward): reverse-time LSTM loops.
//
// Per reverse step the recurrence needs  d x_{i-1} = dgates_i . W   (K = 4D gate rows -> N output columns), the
// transpose of the forward product.  2-D weight-stationary partition: CTA (kb, nb, bh) keeps the bf16 block
// W[K-block kb (4 x UK gate rows), N-block nb] in shared memory for the whole sequence, owns a batch half, and
//   P1  runs the LSTM-cell backward for a 1/NB share of its K-block's hidden units (dgates -> fp32 for the dW GEMMs,
//       bf16 for the tensor cores),
//   --  grid barrier
//   P2  streams the bf16 dgates of its K-block (32 x 4UK) and multiplies (warps split N; ldmatrix.trans B fragments)
//       writing an fp32 partial [32 x UN] that the next step's P1 sums over the KB K-blocks (deterministic order),
//   --  grid barrier.
// Reference semantics: autograd replay of modules/layers.py:18-47 (train.py:83).
#include <cuda_bf16.h>
#include "decoder_internal.cuh"

namespace b200tts {

namespace {

constexpr int PT = 256;
constexpr int BT = 32;
constexpr int KB = 8;            // K-blocks (over hidden units)
constexpr int NBK = 8;           // N-blocks (over output columns)

struct BwdLoopArgs {
    int B, T, D, NOUT, UK, UN, NBH;        // NOUT output columns (D for the generator loop), UN = ceil(NOUT / NBK / 8) * 8
    const float* W; int ldw;               // fp32 [4D, ldw]: dgates . W
    const float* gates;                    // [T, B, 4D] activated gates (forward)
    const float* cstate;                   // [T+1, B, D]
    const float* dh_static;                // [T, B, D]
    const uint8_t* mask_h; const uint8_t* mask_c;
    int kind, training; float rate_h, rate_c;
    float* dgates;                         // [T, B, 4D] out (fp32)
    __nv_bfloat16* dgb;                    // [B, 4D] staging (bf16)
    float* part;                           // [KB, B, NOUT] partial products of the previous reverse step
    int hcol;                              // column of d h inside the NOUT outputs (0 for the generator loop)
    unsigned* barrier; int* abort_flag;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit_wait() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool grid_barrier(unsigned* counter, unsigned& target, unsigned nblocks, int* abort_flag) {
    __shared__ int s_ok;
    __syncthreads();
    if (threadIdx.x == 0) {
        target += nblocks;
        __threadfence();
        atomicAdd(counter, 1u);
        int ok = 1;
        const long long t0 = clock64();
        while (ld_acquire(counter) < target) {
            if (clock64() - t0 > 4000000000ll || *reinterpret_cast<volatile int*>(abort_flag)) { ok = 0; *abort_flag = 1; break; }
        }
        __threadfence();
        s_ok = ok;
    }
    __syncthreads();
    return s_ok != 0;
}

// Generator-LSTM reverse loop (no attention): NOUT = D.
__global__ void __launch_bounds__(PT, 1) lstm_bwd_loop_kernel(const BwdLoopArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x;
    const int kb = cta % KB, nb = (cta / KB) % NBK, bh = cta / (KB * NBK);
    const int B = p.B, D = p.D, UK = p.UK, UN = p.UN, KROWS = 4 * UK;
    const int WLD = UN + 8, ALD = KROWS + 8;
    const int b0 = bh * BT, n0 = nb * UN;
    __nv_bfloat16* Ws = reinterpret_cast<__nv_bfloat16*>(smem_raw);                 // [KROWS][WLD]  (k rows, n contiguous)
    __nv_bfloat16* As = Ws + (size_t)KROWS * WLD;                                    // [BT][ALD]
    const unsigned nblocks = gridDim.x;

    // resident weight block: row r = g*UK + uk  <->  gate row g*D + kb*UK + uk ; column n <-> output n0 + n
    for (int idx = tid; idx < KROWS * UN; idx += PT) {
        const int r = idx / UN, n = idx % UN;
        const int g = r / UK, uk = r % UK;
        float w = 0.f;
        if (n0 + n < p.NOUT) w = p.W[(size_t)(g * D + kb * UK + uk) * p.ldw + n0 + n];
        Ws[r * WLD + n] = __float2bfloat16_rn(w);
    }
    __syncthreads();

    // P1 ownership: hidden units [kb*UK + nb*UP, +UP) with UP = UK / NBK, for the 32 utterances of this batch half
    const int UP = UK / NBK;
    const float inv_h = 1.f / (1.f - p.rate_h), inv_c = 1.f / (1.f - p.rate_c);
    constexpr int MAXE = 4;                       // (b, u) pairs per thread: BT * UP / PT  (UP <= 32)
    float dc_reg[MAXE], dhz_reg[MAXE];
#pragma unroll
    for (int e = 0; e < MAXE; ++e) { dc_reg[e] = 0.f; dhz_reg[e] = 0.f; }
    unsigned target = 0;

    for (int i = p.T - 1; i >= 0; --i) {
        const bool last = (i == p.T - 1);
        // ---------------- P1: LSTM cell backward ----------------
#pragma unroll
        for (int e = 0; e < MAXE; ++e) {
            const int idx = tid + e * PT;
            if (idx < BT * UP) {
                const int bl = idx / UP, up = idx % UP, b = b0 + bl, u = kb * UK + nb * UP + up;
                if (b < B) {
                    const size_t bu = (size_t)b * D + u, g0 = ((size_t)i * B + b) * 4 * D + u;
                    float dh = p.dh_static[(size_t)i * B * D + bu];
                    float dc_in = 0.f;
                    if (!last) {
                        float rec = 0.f;
                        for (int k2 = 0; k2 < KB; ++k2) rec += __ldcg(p.part + ((size_t)k2 * B + b) * p.NOUT + p.hcol + u);
                        dh += rec + dhz_reg[e];
                        dc_in = dc_reg[e];
                    }
                    const float gi = p.gates[g0], gf = p.gates[g0 + D], gg = p.gates[g0 + 2 * D], go = p.gates[g0 + 3 * D];
                    const float cp = p.cstate[(size_t)i * B * D + bu];
                    const float tc = tanhf(gf * cp + gi * gg);
                    const size_t mi = (size_t)i * B * D + bu;
                    float dhn, dcn, dc_prev_direct = 0.f, dh_prev_direct = 0.f;
                    if (p.kind == B200TTS_CELL_ZONEOUT) {
                        float kh, kc;
                        if (p.training) {
                            kh = (1.f - p.rate_h) * (p.mask_h ? (float)p.mask_h[mi] * inv_h : 1.f);
                            kc = (1.f - p.rate_c) * (p.mask_c ? (float)p.mask_c[mi] * inv_c : 1.f);
                        } else { kh = 1.f - p.rate_h; kc = 1.f - p.rate_c; }
                        dhn = dh * kh; dh_prev_direct = dh - dhn;
                        dcn = dc_in * kc + dhn * go * (1.f - tc * tc);
                        dc_prev_direct = dc_in - dc_in * kc;
                    } else {
                        dhn = (p.training && p.mask_h) ? dh * (float)p.mask_h[mi] * inv_h : dh;
                        dcn = dc_in + dhn * go * (1.f - tc * tc);
                    }
                    const float di = dcn * gg * gi * (1.f - gi), df = dcn * cp * gf * (1.f - gf);
                    const float dg = dcn * gi * (1.f - gg * gg), dO = dhn * tc * go * (1.f - go);
                    p.dgates[g0] = di; p.dgates[g0 + D] = df; p.dgates[g0 + 2 * D] = dg; p.dgates[g0 + 3 * D] = dO;
                    __nv_bfloat16* db = p.dgb + (size_t)b * 4 * D + u;
                    db[0] = __float2bfloat16_rn(di); db[D] = __float2bfloat16_rn(df);
                    db[2 * D] = __float2bfloat16_rn(dg); db[3 * D] = __float2bfloat16_rn(dO);
                    dc_reg[e] = dcn * gf + dc_prev_direct;
                    dhz_reg[e] = dh_prev_direct;
                }
            }
        }
        if (!grid_barrier(p.barrier, target, nblocks, p.abort_flag)) return;
        if (i == 0) break;

        // ---------------- P2: partial[kb] = dgates[:, K-block kb] . W[K-block kb, N-block nb] ----------------
        {
            const int segs = UK / 8;                       // 16-byte segments per gate block per row
            for (int idx = tid; idx < BT * 4 * segs; idx += PT) {
                const int r = idx / (4 * segs), rem = idx % (4 * segs), g = rem / segs, sg = rem % segs;
                __nv_bfloat16* d = As + r * ALD + g * UK + sg * 8;
                if (b0 + r < B) cp_async16(d, p.dgb + (size_t)(b0 + r) * 4 * D + g * D + kb * UK + sg * 8);
                else *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);
            }
            cp_async_commit_wait();
            __syncthreads();
            // warps split N: warp w owns n-tile pairs {w, w+8, ...} (16 columns each)
            const int npairs = UN / 16;
            for (int np = warp; np < npairs; np += 8) {
                float acc[2][2][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
                for (int kk = 0; kk < KROWS; kk += 16) {
                    uint32_t af[2][4], bf[4];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
                        ldmatrix_x4(af[mt][0], af[mt][1], af[mt][2], af[mt][3], As + (mt * 16 + (lane & 15)) * ALD + kk + (lane >> 4) * 8);
                    ldmatrix_x4_trans(bf[0], bf[1], bf[2], bf[3],
                                      Ws + (size_t)(kk + (lane & 7) + ((lane >> 3) & 1) * 8) * WLD + np * 16 + (lane >> 4) * 8);
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        mma_bf16(acc[mt][0], af[mt], bf[0], bf[1]);
                        mma_bf16(acc[mt][1], af[mt], bf[2], bf[3]);
                    }
                }
                const int g = lane >> 2, tq = lane & 3;
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int b = b0 + mt * 16 + g + 8 * (e >> 1);
                            const int n = n0 + np * 16 + nt * 8 + 2 * tq + (e & 1);
                            if (b < B && n < p.NOUT) p.part[((size_t)kb * B + b) * p.NOUT + n] = acc[mt][nt][e];
                        }
            }
        }
        if (!grid_barrier(p.barrier, target, nblocks, p.abort_flag)) return;
    }
}

}  // namespace

bool persist_bwd_supported(const b200tts_decoder_shape& s) {
    if (s.D % (KB * 16) != 0 || s.B > 2 * BT) return false;
    const int UK = s.D / KB;
    if (UK / NBK > 32 || (UK % NBK) != 0) return false;
    return true;
}

size_t persist_bwd_gen_extra_bytes(const b200tts_decoder_shape& s) {
    // dgb [B, 4D] bf16 + part [KB, B, D] fp32 + barrier
    return ((size_t)s.B * 4 * s.D * 2 + 255) / 256 * 256 + ((size_t)KB * s.B * s.D * 4 + 255) / 256 * 256 + 256;
}

// dgates for all T steps of the generator LSTM.  `extra` = persist_bwd_gen_extra_bytes scratch.
int persist_gen_bwd_loop(const b200tts_decoder_shape& s, const b200tts_decoder_params& w, const b200tts_decoder_inputs& in,
                         const DecoderLayout& fl, const float* fws, const float* dh_static, float* dgates, unsigned char* extra,
                         cudaStream_t st) {
    const int B = s.B, D = s.D;
    BwdLoopArgs a{};
    a.B = B; a.T = s.T; a.D = D; a.NOUT = D; a.UK = D / KB; a.UN = (cdiv(D, NBK) + 15) / 16 * 16; a.NBH = (B + BT - 1) / BT;
    a.W = w.gen_w_hh; a.ldw = D;
    a.gates = fws + fl.gg; a.cstate = fws + fl.cg; a.dh_static = dh_static;
    a.mask_h = in.mask_gen_h; a.mask_c = in.mask_gen_c; a.kind = s.cell_kind; a.training = s.training; a.rate_h = s.rate_h; a.rate_c = s.rate_c;
    a.dgates = dgates;
    size_t off = 0;
    a.dgb = reinterpret_cast<__nv_bfloat16*>(extra + off); off += ((size_t)B * 4 * D * 2 + 255) / 256 * 256;
    a.part = reinterpret_cast<float*>(extra + off); off += ((size_t)KB * B * D * 4 + 255) / 256 * 256;
    a.barrier = reinterpret_cast<unsigned*>(extra + off);
    a.abort_flag = reinterpret_cast<int*>(a.barrier + 32);
    a.hcol = 0;
    B200_CUDA(cudaMemsetAsync(a.barrier, 0, 256, st));
    const size_t smem = ((size_t)4 * a.UK * (a.UN + 8) + (size_t)BT * (4 * a.UK + 8)) * 2;
    B200_REQUIRE(smem <= 227 * 1024, "persistent backward: %zu B of shared memory needed", smem);
    void* fn = (void*)lstm_bwd_loop_kernel;
    B200_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = KB * NBK * a.NBH;
    int per_sm = 0, dev = 0, sms = 0;
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, PT, smem));
    B200_CUDA(cudaGetDevice(&dev));
    B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    B200_REQUIRE(per_sm * sms >= grid, "persistent backward: %d CTAs cannot be co-resident", grid);
    void* params[] = {&a};
    B200_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(PT), params, smem, st));
    B200_LAUNCH_CHECK();
    return B200TTS_OK;
}

}  // namespace b200tts
