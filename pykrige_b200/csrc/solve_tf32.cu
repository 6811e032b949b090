// solve_tf32.cu — K3 for dtype = KB200_F32: the same covariance-form contraction q_j = ||W c_j||^2
//  on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the accumulator in
// TMEM. fp32 accuracy comes from the 3xTF32 split  W = Wh + Wl, c = ch + cl (each part exactly
// representable in TF32):  W c ~= Wh ch + Wh cl + Wl ch  accumulated in fp32 (measured error of
// sigma^2: 3e-6 relative, tolerance for fp32 is 1e-2).
//
// Orientation: D[point][row of W] so that one TMEM lane = one prediction point; the epilogue then
// needs no cross-lane reduction (each thread squares-and-adds its own 256 columns).
//   A operand (M = 128 points)  : RHS tile, K-major, from the per-CTA scratch ring (generated once per
//                                 point tile by all warps, fp32 sqrt/exp on fp64 coordinate differences)
//   B operand (N = 256 W rows)  : W tile, K-major, packed by pack_tf32_kernel
//   both in the canonical no-swizzle ("interleaved") UMMA layout: 8-row x 16-byte core matrices,
//   k-chunks 128 B apart (LBO), 8-row groups 512 B apart (SBO); one stage = 16 k = 2 MMA k-steps.
// Roles: warp 0 = bulk-copy producer (cp.async.bulk + mbarrier), warp 1 = MMA issuer (one lane),
// warp 2 = TMEM allocator, warps 4..7 = epilogue (tcgen05.ld) + per-point finalize.
// Two 256-column accumulators alternate between row blocks so the epilogue of block I overlaps the
// MMAs of block I+1.
#include "common.cuh"
#include "kernels.h"

#define TF_STAGES 4
#define TF_THREADS 512
#define TF_GEN_THREADS 256              // warps 8..15: two generator threads per prediction point
#define TF_TM 128                  // points per CTA tile (UMMA M)
#define TF_BN KB_BM                // W rows per row block (UMMA N) = 256
#define TF_BK KB_BK                // k per stage = 16
#define TF_W_BYTES (TF_BN * TF_BK * 4 * 2)     // hi + lo = 32 KB
#define TF_C_BYTES (TF_TM * TF_BK * 4 * 2)     // hi + lo = 16 KB
#define TF_STAGE_BYTES (TF_W_BYTES + TF_C_BYTES)

__device__ __forceinline__ uint32_t tf_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tf_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(tf_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tf_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(tf_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tf_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(tf_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tf_mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t it = 0; it < (1u << 26); ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(tf_smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();     // a lost arrival traps instead of hanging the GPU
}
__device__ __forceinline__ void tf_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(tf_smem_u32(dst)), "l"(src), "r"(bytes), "r"(tf_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float tf32_round(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
// K-major, no swizzle: start address, LBO (k-chunk stride) = 128 B, SBO (8-row group stride) = 512 B,
// descriptor version 1 (Blackwell), layout type 0 (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t tf_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(512u >> 4) << 32) |
           (1ull << 46);
}
// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), K-major A and B,
// N = 256 (bits 17-22 = N >> 3), M = 128 (bits 24-28 = M >> 4)
#define TF_IDESC ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TF_BN >> 3) << 17) | ((uint32_t)(TF_TM >> 4) << 24))

__device__ __forceinline__ void tf_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(TF_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tf_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                 :: "r"(tf_smem_u32(bar)) : "memory");
}

// ---- pack: W (fp64, row-major lower triangle) + dual rows -> TF32 hi/lo tiles in UMMA layout --------
// tile (row block I, k stage t): 8192 floats = [hi 4096][lo 4096];
//   element (r, k) of a part at float offset (r/8)*128 + (k/4)*32 + (r%8)*4 + (k%4)
__global__ void __launch_bounds__(256) pack_tf32_kernel(const double* __restrict__ W, int ld, int n, int n_pad, int na,
                                                         const double* __restrict__ Uz, PackMap pm,
                                                         float* __restrict__ out) {
    int I = blockIdx.y, kt = blockIdx.x;
    if (kt >= pm.ktiles[I]) return;
    float* o = out + ((size_t)pm.tile_off[I] + kt) * (TF_W_BYTES / 4);
    for (int e = threadIdx.x; e < TF_BN * TF_BK; e += 256) {
        int rg = e >> 7, kc = (e >> 5) & 3, rr = (e >> 2) & 7, kk = e & 3;
        int r = I * TF_BN + rg * 8 + rr;
        int k = kt * TF_BK + kc * 4 + kk;
        double v = 0.0;
        if (r < n) { if (k <= r) v = W[(size_t)r * ld + k]; }
        else if (r < n + na) { if (k < n) v = Uz[(size_t)(r - n) * n_pad + k]; }
        float hi = tf32_round((float)v);
        float lo = tf32_round((float)(v - (double)hi));
        o[e] = hi;
        o[TF_BN * TF_BK + e] = lo;
    }
}

template <int DIM, int MODEL>
__device__ __forceinline__ float tf_cov_rhs(const VgParams& v, double dd) {
    // exact hit on the fp64 distance (|d| <= eps, ok.py:665-672); the variogram itself in fp32
    if (v.exact && dd <= v.eps) return (float)v.c0;
    float d = (float)dd;
    float c0 = (float)v.c0, p0 = (float)v.p0, p1 = (float)v.p1, p2 = (float)v.p2;
    float g;
    if (MODEL == KB200_VG_LINEAR) g = p0 * d + p1;
    else if (MODEL == KB200_VG_POWER) g = p0 * powf(d, p1) + p2;
    else if (MODEL == KB200_VG_GAUSSIAN) { float r = p1 * (4.0f / 7.0f); g = p0 * (1.0f - expf(-(d * d) / (r * r))) + p2; }
    else if (MODEL == KB200_VG_EXPONENTIAL) g = p0 * (1.0f - expf(-d / (p1 / 3.0f))) + p2;
    else if (MODEL == KB200_VG_SPHERICAL) {
        if (d <= p1) { float q = d / p1; g = p0 * (1.5f * q - 0.5f * q * q * q) + p2; } else g = p0 + p2;
    } else if (MODEL == KB200_VG_TABLE) g = (float)kb_gamma<KB200_VG_TABLE>(v, dd);     // tabulated callable (fp64 table)
    else { float q = d / (p1 / 3.0f); g = p0 * (1.0f - (1.0f - q) * expf(-q)) + p2; }
    return c0 - g;
}

template <int DIM, int MODEL>
__global__ void __launch_bounds__(TF_THREADS, 1) solve_kernel_tf32(const __grid_constant__ SolvePtParams P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* stage_base = smem_raw;                                            // TF_STAGES * 48 KB
    float* auxs = reinterpret_cast<float*>(smem_raw + (size_t)TF_STAGES * TF_STAGE_BYTES);   // KB_MAXAUX * 128
    uint64_t* full = reinterpret_cast<uint64_t*>(auxs + KB_MAXAUX * TF_TM);          // TF_STAGES
    uint64_t* empty = full + TF_STAGES;                                              // TF_STAGES
    uint64_t* tfull = empty + TF_STAGES;                                             // 2
    uint64_t* tempty = tfull + 2;                                                    // 2
    uint64_t* gfull = tempty + 2;                                                    // 2
    uint64_t* gempty = gfull + 2;                                                    // 2
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(gempty + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nk = (P.n + TF_BK - 1) / TF_BK;
    const size_t sbuf = (size_t)nk * TF_C_BYTES;                                     // one RHS column block
    unsigned char* scratch = reinterpret_cast<unsigned char*>(P.scratch) + (size_t)blockIdx.x * 2 * sbuf;
    const unsigned char* gt = reinterpret_cast<const unsigned char*>(P.tiles);
    const long long ntiles = (P.m + TF_TM - 1) / TF_TM;

    if (tid == 0) {
        for (int s = 0; s < TF_STAGES; ++s) { tf_mbar_init(&full[s], 1); tf_mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { tf_mbar_init(&tfull[b], 1); tf_mbar_init(&tempty[b], 4); }
        for (int b = 0; b < 2; ++b) { tf_mbar_init(&gfull[b], TF_GEN_THREADS); tf_mbar_init(&gempty[b], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     :: "r"(tf_smem_u32(tmem_base_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *tmem_base_smem;

    // Warp roles (512 threads): warp 0 lane 0 = bulk-copy producer, warp 1 lane 0 = MMA issuer, warp 2 = TMEM allocator,
    // warps 4-7 = epilogue, warps 8-15 = RHS generators working ONE POINT TILE AHEAD of the tensor pipe into the other
    // half of a double-buffered scratch ring (gfull / gempty mbarriers), as in solve_i8.cu.
    if (warp >= 8) {
        const int gt_ = tid - 8 * 32;
        const int pl = gt_ & (TF_TM - 1);
        const int ks = gt_ >> 7;                           // 0..1
        uint32_t it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int b = (int)(it & 1);
            tf_mbar_wait(&gempty[b], ((it >> 1) & 1) ^ 1);
            unsigned char* sc = scratch + (size_t)b * sbuf;
            const long long pj = tile * TF_TM + pl;
            const bool pvalid = pj < P.m;
            double px = 0.0, py = 0.0, pz = 0.0;
            if (pvalid) kb_load_point<DIM>(P.ps, P.an, pj, px, py, pz);
            for (int t = ks; t < nk; t += TF_GEN_THREADS / TF_TM) {
                float* ct = reinterpret_cast<float*>(sc + (size_t)t * TF_C_BYTES);
#pragma unroll
                for (int kc = 0; kc < 4; ++kc) {
                    float hi[4], lo[4];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int k = t * TF_BK + kc * 4 + kk;
                        float c = 0.0f;
                        if (pvalid && k < P.n) {
                            double dd = kb_dist<DIM>(__ldg(P.ax + k), __ldg(P.ay + k), KB_HASZ(DIM) ? __ldg(P.az + k) : 0.0,
                                                     px, py, pz);
                            c = tf_cov_rhs<DIM, MODEL>(P.vg, dd);
                        }
                        hi[kk] = tf32_round(c);
                        lo[kk] = tf32_round(c - hi[kk]);
                    }
                    const int off = (pl >> 3) * 128 + kc * 32 + (pl & 7) * 4;      // floats
                    *reinterpret_cast<float4*>(ct + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(ct + TF_TM * TF_BK + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            __threadfence();
            asm volatile("fence.proxy.async.global;\n" ::: "memory");
            tf_mbar_arrive(&gfull[b]);
        }
    } else if (warp == 0) {
        if (lane == 0) {
            const uint64_t pol_w = kb_policy_evict_last(), pol_c = kb_policy_evict_first();
            uint32_t gg = 0, it = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int b = (int)(it & 1);
                const unsigned char* sc = scratch + (size_t)b * sbuf;
                tf_mbar_wait(&gfull[b], (it >> 1) & 1);
                long long tau = 0;
                for (int I = 0; I < P.nrb; ++I) {
                    const int kt = P.pm.ktiles[I];
                    for (int t = 0; t < kt; ++t, ++tau, ++gg) {
                        const int s = gg % TF_STAGES;
                        tf_mbar_wait(&empty[s], (uint32_t)(((gg / TF_STAGES) & 1) ^ 1));
                        tf_mbar_expect_tx(&full[s], TF_STAGE_BYTES);
                        unsigned char* sb = stage_base + (size_t)s * TF_STAGE_BYTES;
                        kb_bulk_g2s_hint(sb, gt + (size_t)tau * TF_W_BYTES, TF_W_BYTES, &full[s], pol_w);
                        kb_bulk_g2s_hint(sb + TF_W_BYTES, sc + (size_t)t * TF_C_BYTES, TF_C_BYTES, &full[s], pol_c);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t gg = 0, gb = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int I = 0; I < P.nrb; ++I, ++gb) {
                    const int kt = P.pm.ktiles[I];
                    const int buf = gb & 1;
                    tf_mbar_wait(&tempty[buf], (uint32_t)(((gb >> 1) & 1) ^ 1));
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint32_t d_tmem = tmem_base + (uint32_t)buf * TF_BN;
                    for (int t = 0; t < kt; ++t, ++gg) {
                        const int s = gg % TF_STAGES;
                        tf_mbar_wait(&full[s], (uint32_t)((gg / TF_STAGES) & 1));
                        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                        const uint32_t wb = tf_smem_u32(stage_base + (size_t)s * TF_STAGE_BYTES);
                        const uint32_t w_hi = wb, w_lo = wb + TF_W_BYTES / 2;
                        const uint32_t c_hi = wb + TF_W_BYTES, c_lo = c_hi + TF_C_BYTES / 2;
#pragma unroll
                        for (int kstep = 0; kstep < TF_BK / 8; ++kstep) {
                            const uint32_t ko = (uint32_t)kstep * 256u;          // 2 k-chunks of 128 B
                            const uint32_t first = (t == 0 && kstep == 0) ? 0u : 1u;
                            tf_mma(d_tmem, tf_desc(c_hi + ko), tf_desc(w_hi + ko), first);
                            tf_mma(d_tmem, tf_desc(c_hi + ko), tf_desc(w_lo + ko), 1u);
                            tf_mma(d_tmem, tf_desc(c_lo + ko), tf_desc(w_hi + ko), 1u);
                        }
                        tf_commit(&empty[s]);              // stage free once these MMAs have read it
                    }
                    tf_commit(&tfull[buf]);                // accumulator of row block I complete
                }
            }
        }
    } else if (warp >= 4) {
        // epilogue: thread = TMEM lane = prediction point
        const int pl = (warp & 3) * 32 + lane;
        uint32_t gb = 0, it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            double q = 0.0;
            for (int I = 0; I < P.nrb; ++I, ++gb) {
                const int buf = gb & 1;
                tf_mbar_wait(&tfull[buf], (uint32_t)((gb >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t t_addr = tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)buf * TF_BN;
#pragma unroll 1
                for (int ch = 0; ch < TF_BN / 32; ++ch) {
                    uint32_t v[32];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(t_addr + (uint32_t)ch * 32u));
                    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                    const int r0 = I * TF_BN + ch * 32;
                    if (r0 + 31 < P.n) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { double x = (double)__uint_as_float(v[j]); q += x * x; }
                    } else if (r0 < P.n + P.na) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int r = r0 + j;
                            float x = __uint_as_float(v[j]);
                            if (r < P.n) q += (double)x * (double)x;
                            else if (r < P.n + P.na) auxs[(r - P.n) * TF_TM + pl] = x;
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                __syncwarp();
                if (lane == 0) tf_mbar_arrive(&tempty[buf]);
            }
            // ---------------- phase F: finalize, thread = point ----------------
            const long long pj = tile * TF_TM + pl;
            if (pj < P.m) kb_finalize_point<DIM, float>(P, pj, q, auxs + pl, TF_TM);
            __syncwarp();
            if (lane == 0) tf_mbar_arrive(&gempty[(int)(it & 1)]);      // this tile's scratch half may be rewritten
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------------------
static size_t tf32_smem() {
    return (size_t)TF_STAGES * TF_STAGE_BYTES + (size_t)KB_MAXAUX * TF_TM * sizeof(float) +
           (2 * TF_STAGES + 8) * sizeof(uint64_t) + 64;
}

size_t kbk_solve_tf32_scratch_bytes(int n, int grid) {
    return (size_t)grid * 2 * ((n + TF_BK - 1) / TF_BK) * TF_C_BYTES;      // double-buffered
}
int kbk_solve_tf32_tile_points() { return TF_TM; }

template <int DIM, int MODEL>
static cudaError_t tf32_attr() {
    return cudaFuncSetAttribute(solve_kernel_tf32<DIM, MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tf32_smem());
}

cudaError_t kbk_solve_tf32_init() {
#define KB_ATTR(M) KB_CUDA_OK((tf32_attr<2, M>())); KB_CUDA_OK((tf32_attr<3, M>())); KB_CUDA_OK((tf32_attr<KB_GEO, M>()));
    KB_ATTR(KB200_VG_LINEAR) KB_ATTR(KB200_VG_POWER) KB_ATTR(KB200_VG_GAUSSIAN)
    KB_ATTR(KB200_VG_EXPONENTIAL) KB_ATTR(KB200_VG_SPHERICAL) KB_ATTR(KB200_VG_HOLE_EFFECT) KB_ATTR(KB200_VG_TABLE)
#undef KB_ATTR
    return cudaSuccess;
}

template <int DIM>
static cudaError_t tf32_dim(const SolvePtParams& p, int grid, cudaStream_t st) {
    size_t sm = tf32_smem();
    switch (p.vg.model) {
#define KB_CASE(M) case M: solve_kernel_tf32<DIM, M><<<grid, TF_THREADS, sm, st>>>(p); break;
        KB_CASE(KB200_VG_LINEAR) KB_CASE(KB200_VG_POWER) KB_CASE(KB200_VG_GAUSSIAN)
        KB_CASE(KB200_VG_EXPONENTIAL) KB_CASE(KB200_VG_SPHERICAL) KB_CASE(KB200_VG_HOLE_EFFECT) KB_CASE(KB200_VG_TABLE)
#undef KB_CASE
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t kbk_solve_tf32(int dim, const SolvePtParams& p, int grid, cudaStream_t st) {
    if (dim == KB_GEO) return tf32_dim<KB_GEO>(p, grid, st);
    return dim == 2 ? tf32_dim<2>(p, grid, st) : tf32_dim<3>(p, grid, st);
}

cudaError_t kbk_pack_tf32(const double* W, int ld, int n, int n_pad, int na, const double* Uz, const PackMap& pm,
                          void* out, cudaStream_t st) {
    int maxkt = 0;
    for (int i = 0; i < pm.nrb; ++i) maxkt = pm.ktiles[i] > maxkt ? pm.ktiles[i] : maxkt;
    dim3 grid(maxkt, pm.nrb);
    pack_tf32_kernel<<<grid, 256, 0, st>>>(W, ld, n, n_pad, na, Uz, pm, (float*)out);
    return cudaGetLastError();
}
