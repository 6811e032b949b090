// solve.cu — K3: the per-grid-point hot kernel of the global kriging path, and the
// tiny per-point finalize.
//
// Reference semantics (ok.py:665-681, uk.py:942-1007): for every prediction point j
//     b_j = [-gamma(|p_j - x_k|) (0 on an exact hit) ; drift(p_j) ; 1],  x_j = A^-1 b_j,
//     z_j = x_j[:n] . Z,   sigma2_j = -x_j . b_j .
// Here, with c_j = c0 + b_j[:n] and W = chol(C)^-1:
//     q_j  = || W c_j ||^2            <- the O(n^2) contraction, a triangular GEMM on the fp64 tensor pipe
//     g_j  = U^T c_j, zc_j = zeta . c_j   <- extra dense "dual rows" appended to W (same GEMM)
//     finalize: r = g - f_j, mu = S^-1 r, sigma2 = c0 - q + r.mu, z = zc - mu.phi
//
// Kernels in this file (all fp64, mma.sync.m8n8k4 = SASS DMMA):
//   solve_kernel_pt   (K3 v3, the product path) persistent CTAs, one CTA = 64 points x all rows of W, RHS column
//                     block generated once per tile, W/RHS tiles streamed by cp.async.bulk + mbarrier, fused
//                     finalize. See the comment above the kernel.
//   solve_kernel_f64 + finalize_kernel (K3 v1) CTA = (row block x point tile), RHS regenerated per row block;
//                     kept behind KB200_SOLVE_V1=1 for A/B profiling only.
// The tcgen05 variants live in solve_tf32.cu (dtype float32) and solve_i8.cu (dtype float64x).
#include "common.cuh"
#include "kernels.h"
#include <cstdlib>

#define SV_STAGES 3
#define SV_THREADS 256

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // bounded spin: a lost transaction traps instead of hanging the GPU
    for (uint32_t it = 0; it < (1u << 26); ++it)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}
// 1-D bulk copy global -> shared through the TMA engine (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int DIM, int MODEL>
__global__ void __launch_bounds__(SV_THREADS, 1) solve_kernel_f64(const __grid_constant__ SolveParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* Ts = reinterpret_cast<double*>(smem_raw);                   // SV_STAGES * BM*BK
    double* Bs = Ts + SV_STAGES * KB_BM * KB_BK;                        // SV_STAGES * BK*TN
    double* red = Bs + SV_STAGES * KB_BK * KB_TN;                       // 8 * 64
    uint64_t* full = reinterpret_cast<uint64_t*>(red + 8 * KB_TN);      // SV_STAGES

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int jt = blockIdx.x;
    const int I = P.nrb - 1 - (int)blockIdx.y;          // heavy row blocks first
    const int nkt = P.pm.ktiles[I];
    const double* gt = reinterpret_cast<const double*>(P.tiles) + (size_t)P.pm.tile_off[I] * (KB_BM * KB_BK);
    constexpr uint32_t TILE_BYTES = KB_BM * KB_BK * sizeof(double);

    // the prediction point this thread generates RHS entries for
    const long long pj = (long long)jt * KB_TN + warp * 8 + (lane >> 2);
    const bool pvalid = pj < P.m;
    double px = 0.0, py = 0.0, pz = 0.0;
    if (pvalid) kb_load_point<DIM>(P.ps, P.an, pj, px, py, pz);

    if (tid == 0) {
        for (int s = 0; s < SV_STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    // which k tiles this warp's 32 rows need: W rows are lower-triangular, dual rows are dense
    const int r0w = I * KB_BM + warp * 32;
    int warp_kmax;
    if (r0w + 31 >= P.n && r0w < P.n + P.na) warp_kmax = 0x7fffffff;
    else if (r0w >= P.n + P.na) warp_kmax = -1;
    else warp_kmax = r0w + 31;

    auto issue_load = [&](int t) {
        int s = t % SV_STAGES;
        mbar_expect_tx(&full[s], TILE_BYTES);
        bulk_g2s(Ts + (size_t)s * KB_BM * KB_BK, gt + (size_t)t * KB_BM * KB_BK, TILE_BYTES, &full[s]);
    };
    auto gen_rhs = [&](int t) {
        double* bs = Bs + (size_t)(t % SV_STAGES) * KB_BK * KB_TN;
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            int k = t * KB_BK + 4 * s4 + (lane & 3);
            double v = 0.0;
            if (pvalid && k < P.n) {
                double d = kb_dist<DIM>(P.ax[k], P.ay[k], KB_HASZ(DIM) ? P.az[k] : 0.0, px, py, pz);
                v = kb_cov_rhs<MODEL>(P.vg, d);
            }
            bs[(s4 * 8 + warp) * 32 + lane] = v;
        }
    };

    double acc[4][8][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }

    for (int t = 0; t < SV_STAGES - 1 && t < nkt; ++t) {
        if (tid == 0) issue_load(t);
        gen_rhs(t);
    }
    for (int it = 0; it < nkt; ++it) {
        const int s = it % SV_STAGES;
        mbar_wait(&full[s], (uint32_t)((it / SV_STAGES) & 1));
        __syncthreads();      // everyone is done with tile it-1 (its stage is refilled below); RHS tile `it` is visible
        const int tn = it + SV_STAGES - 1;
        if (tn < nkt) {
            if (tid == 0) issue_load(tn);
            gen_rhs(tn);
        }
        if (it * KB_BK <= warp_kmax) {
            const double* ts = Ts + (size_t)s * KB_BM * KB_BK;
            const double* bs = Bs + (size_t)s * KB_BK * KB_TN;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                double fa[4], fb[8];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) fa[mt] = ts[(k4 * 32 + warp * 4 + mt) * 32 + lane];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) fb[nt] = bs[(k4 * 8 + nt) * 32 + lane];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt)
                        kb_dmma(acc[mt][nt][0], acc[mt][nt][1], fa[mt], fb[nt]);
            }
        }
    }

    // epilogue: W rows -> sum of squares per point; dual rows -> direct output
    double qs[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { qs[nt][0] = 0.0; qs[nt][1] = 0.0; }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int r = r0w + mt * 8 + (lane >> 2);
        if (r < P.n) {
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                qs[nt][0] += acc[mt][nt][0] * acc[mt][nt][0];
                qs[nt][1] += acc[mt][nt][1] * acc[mt][nt][1];
            }
        } else if (r < P.n + P.na) {
            double* ao = P.auxout + (size_t)(r - P.n) * P.mpad + (size_t)jt * KB_TN + 2 * (lane & 3);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                ao[nt * 8] = acc[mt][nt][0];
                ao[nt * 8 + 1] = acc[mt][nt][1];
            }
        }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            double v = qs[nt][i];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            qs[nt][i] = v;
        }
    if ((lane >> 2) == 0) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            red[warp * KB_TN + nt * 8 + 2 * lane] = qs[nt][0];
            red[warp * KB_TN + nt * 8 + 2 * lane + 1] = qs[nt][1];
        }
    }
    __syncthreads();
    if (tid < KB_TN) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w * KB_TN + tid];     // fixed order: deterministic
        P.partial[(size_t)I * P.mpad + (size_t)jt * KB_TN + tid] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// K3 v3: persistent, warp-specialised, one CTA = one tile of 64 prediction points x ALL rows of W.
//
// v1 (one CTA per (row block, point tile)) regenerates every RHS tile once per row block - a 10x
// recompute of sqrt+exp at N=5000 - and alternates generation with the tensor work (measured: DMMA pipe
// 54 % active). v2 (warp-specialised but still regenerating) was slower: four producer warps cannot
// hide the fp64 sqrt/exp latency. v3 removes the recompute instead:
//   phase G  all 12 warps evaluate the RHS column block c[k][j] of the tile ONCE (n x 64 values) and
//            park it, already in MMA-fragment order, in a per-CTA scratch ring (L2-resident, re-used for
//            every tile the CTA processes; size independent of M);
//   phase M  warp 0 streams W tiles (32 KB) and RHS tiles (8 KB) with cp.async.bulk + mbarrier into a
//            4-stage ring; warps 4..11 do nothing but LDS + DMMA, walking all row blocks and keeping the
//            per-point sum of squares in registers;
//   phase F  the (K+1)x(K+1) drift solve and the two outputs per point are produced in the same CTA:
//            no partial buffers, no separate finalize pass, fixed summation order (deterministic).
#define PT_STAGES 5
#define PT_THREADS 384
#define PT_STAGE_BYTES ((KB_BM * KB_BK + KB_BK * KB_TN) * 8)

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}

template <int DIM, int MODEL>
__global__ void __launch_bounds__(PT_THREADS, 1) solve_kernel_pt(const __grid_constant__ SolvePtParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* Ts = reinterpret_cast<double*>(smem_raw);                   // PT_STAGES * BM*BK
    double* Bs = Ts + PT_STAGES * KB_BM * KB_BK;                        // PT_STAGES * BK*TN
    double* qred = Bs + PT_STAGES * KB_BK * KB_TN;                      // 8 * 64
    double* auxs = qred + 8 * KB_TN;                                    // KB_MAXAUX * 64
    uint64_t* full = reinterpret_cast<uint64_t*>(auxs + KB_MAXAUX * KB_TN);   // PT_STAGES
    uint64_t* empty = full + PT_STAGES;                                 // PT_STAGES

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nk = (P.n + KB_BK - 1) / KB_BK;                           // k tiles of a full column block
    double* scratch = P.scratch + (size_t)blockIdx.x * nk * (KB_BK * KB_TN);
    const double* gt = reinterpret_cast<const double*>(P.tiles);
    const long long ntiles = (P.m + KB_TN - 1) / KB_TN;
    const int K = P.n_rl + P.n_hd, K1 = K + 1;

    if (tid == 0) {
        for (int s = 0; s < PT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    uint32_t git = 0;                 // running stage counter (same sequence in producer and consumers)
    const int cw = warp - 4;          // consumer warp 0..7

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ---------------- phase G: RHS column block of this tile, once ----------------
        {
            const int pl = tid & 63;                 // point within the tile
            const int ks = tid >> 6;                 // 0..5: k-tile slice
            const long long pj = tile * KB_TN + pl;
            const bool pvalid = pj < P.m;
            double px = 0.0, py = 0.0, pz = 0.0;
            if (pvalid) kb_load_point<DIM>(P.ps, P.an, pj, px, py, pz);
            for (int t = ks; t < nk; t += PT_THREADS / 64) {
                double* bt = scratch + (size_t)t * (KB_BK * KB_TN);
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    double v[4];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int k = t * KB_BK + k4 * 4 + kk;
                        double val = 0.0;
                        if (pvalid && k < P.n) {
                            double d = kb_dist<DIM>(__ldg(P.ax + k), __ldg(P.ay + k), KB_HASZ(DIM) ? __ldg(P.az + k) : 0.0,
                                                    px, py, pz);
                            val = kb_cov_rhs<MODEL>(P.vg, d);
                        }
                        v[kk] = val;
                    }
                    // fragment order ((k4*8 + n/8)*32 + (n%8)*4 + k%4): 4 consecutive k = 32 contiguous bytes
                    double2* dst = reinterpret_cast<double2*>(bt + (k4 * 8 + (pl >> 3)) * 32 + (pl & 7) * 4);
                    dst[0] = make_double2(v[0], v[1]);
                    dst[1] = make_double2(v[2], v[3]);
                }
            }
            // generic-proxy global writes -> later read by the async proxy (bulk copies) of this CTA
            __threadfence();
            asm volatile("fence.proxy.async.global;\n" ::: "memory");
        }
        __syncthreads();

        // ---------------- phase M ----------------
        if (warp == 0) {
            if (lane == 0) {
                uint32_t g = git;
                long long tau = 0;                   // W tiles are stored contiguously in (I, t) order
                for (int I = 0; I < P.nrb; ++I) {
                    const int kt = P.pm.ktiles[I];
                    for (int t = 0; t < kt; ++t, ++tau, ++g) {
                        const int s = g % PT_STAGES;
                        mbar_wait(&empty[s], (uint32_t)(((g / PT_STAGES) & 1) ^ 1));
                        mbar_expect_tx(&full[s], PT_STAGE_BYTES);
                        bulk_g2s(Ts + (size_t)s * KB_BM * KB_BK, gt + (size_t)tau * (KB_BM * KB_BK),
                                 KB_BM * KB_BK * 8, &full[s]);
                        bulk_g2s(Bs + (size_t)s * KB_BK * KB_TN, scratch + (size_t)t * (KB_BK * KB_TN),
                                 KB_BK * KB_TN * 8, &full[s]);
                    }
                }
            }
        } else if (warp >= 4) {
            uint32_t g = git;
            double qs[8][2];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) { qs[nt][0] = 0.0; qs[nt][1] = 0.0; }
            for (int I = 0; I < P.nrb; ++I) {
                const int kt = P.pm.ktiles[I];
                // this warp owns the m-tiles cw, cw+8, cw+16, cw+24 of the 256-row block (8 rows each):
                // interleaving keeps the eight warps equally busy inside the triangular diagonal block and
                // lets every m-tile skip the k tiles above the diagonal (W rows are lower-triangular, the
                // dual rows are dense, padding rows need nothing)
                int kmax[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r0 = I * KB_BM + (cw + 8 * q) * 8;
                    if (r0 + 7 >= P.n && r0 < P.n + P.na) kmax[q] = 0x7fffffff;
                    else if (r0 >= P.n + P.na) kmax[q] = -1;
                    else kmax[q] = r0 + 7;
                }
                double acc[4][8][2];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 8; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
                for (int t = 0; t < kt; ++t, ++g) {
                    const int s = g % PT_STAGES;
                    // every consumer waits for every stage (also the ones it skips) so that no warp can lap
                    // the ring and arrive twice on empty[s] within one phase
                    mbar_wait(&full[s], (uint32_t)((g / PT_STAGES) & 1));
                    const int k0 = t * KB_BK;
                    if (k0 <= kmax[3] || k0 <= kmax[2] || k0 <= kmax[1] || k0 <= kmax[0]) {
                        const double* ts = Ts + (size_t)s * KB_BM * KB_BK;
                        const double* bs = Bs + (size_t)s * KB_BK * KB_TN;
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            double fb[8];
#pragma unroll
                            for (int nt = 0; nt < 8; ++nt) fb[nt] = bs[(k4 * 8 + nt) * 32 + lane];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (k0 <= kmax[q]) {
                                    const double fa = ts[(k4 * 32 + cw + 8 * q) * 32 + lane];
#pragma unroll
                                    for (int nt = 0; nt < 8; ++nt)
                                        kb_dmma(acc[q][nt][0], acc[q][nt][1], fa, fb[nt]);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s]);
                }
                // row-block epilogue: W rows -> running sum of squares; dual rows -> shared memory
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) {
                    const int r = I * KB_BM + (cw + 8 * mt) * 8 + (lane >> 2);
                    if (r < P.n) {
                        if (!P.gform) {
#pragma unroll
                            for (int nt = 0; nt < 8; ++nt) {
                                qs[nt][0] += acc[mt][nt][0] * acc[mt][nt][0];
                                qs[nt][1] += acc[mt][nt][1] * acc[mt][nt][1];
                            }
                        } else {
                            // quadratic form c^T G c: multiply row r of T c by c[r] (read back from the
                            // scratch ring: tile r/16, fragment order)
                            const double* bt = scratch + (size_t)(r >> 4) * (KB_BK * KB_TN) + ((r & 15) >> 2) * 256 + (r & 3);
#pragma unroll
                            for (int nt = 0; nt < 8; ++nt) {
                                const int c0i = nt * 8 + 2 * (lane & 3);
                                qs[nt][0] += acc[mt][nt][0] * bt[(c0i >> 3) * 32 + (c0i & 7) * 4];
                                qs[nt][1] += acc[mt][nt][1] * bt[((c0i + 1) >> 3) * 32 + ((c0i + 1) & 7) * 4];
                            }
                        }
                    } else if (r < P.n + P.na) {
                        double* ao = auxs + (r - P.n) * KB_TN + 2 * (lane & 3);
#pragma unroll
                        for (int nt = 0; nt < 8; ++nt) {
                            ao[nt * 8] = acc[mt][nt][0];
                            ao[nt * 8 + 1] = acc[mt][nt][1];
                        }
                    }
                }
            }
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    double v = qs[nt][i];
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    qs[nt][i] = v;
                }
            if ((lane >> 2) == 0) {
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    qred[cw * KB_TN + nt * 8 + 2 * lane] = qs[nt][0];
                    qred[cw * KB_TN + nt * 8 + 2 * lane + 1] = qs[nt][1];
                }
            }
        }
        // every role advances the ring counter by the same amount
        for (int I = 0; I < P.nrb; ++I) git += (uint32_t)P.pm.ktiles[I];
        __syncthreads();

        // ---------------- phase F: per-point finalize ----------------
        if (tid < KB_TN) {
            const long long pj = tile * KB_TN + tid;
            if (pj < P.m) {
                double q = 0.0;
#pragma unroll
                for (int w = 0; w < 8; ++w) q += qred[w * KB_TN + tid];     // fixed order: deterministic
                double r[KB200_MAX_DRIFT + 1];
                double f[KB200_MAX_DRIFT + 1];
                if (P.n_rl > 0) {
                    double x, y, z;
                    kb_load_point<DIM>(P.ps, P.an, pj, x, y, z);
                    f[0] = (x - P.ds.shift[0]) * P.ds.scale[0];
                    f[1] = (y - P.ds.shift[1]) * P.ds.scale[1];
                    if (DIM == 3) f[2] = (z - P.ds.shift[2]) * P.ds.scale[2];
                }
                for (int c = 0; c < P.n_hd; ++c) {
                    double v = P.drift_pts[(size_t)c * P.drift_stride + P.drift_first + pj];
                    f[P.n_rl + c] = (v - P.ds.shift[P.n_rl + c]) * P.ds.scale[P.n_rl + c];
                }
                f[K] = 1.0;
                const double zc = auxs[K1 * KB_TN + tid];
                const double* Sinv = P.consts;
                const double* phi = P.consts + K1 * K1;
                if (P.gform == 2) {
                    // pseudo-inverse form (pinv.cu): b = [c; f], sigma^2 = -b^T A^+ b, z = w1.c + w2.f with
                    // q = c^T G11 c, aux rows = G21 c, consts = G22 | w2
                    double acc = q, zz = zc;
                    for (int a = 0; a < K1; ++a) {
                        double gf = 0.0;
                        for (int b = 0; b < K1; ++b) gf += Sinv[a * K1 + b] * f[b];
                        acc += f[a] * (2.0 * auxs[a * KB_TN + tid] + gf);
                        zz += phi[a] * f[a];
                    }
                    P.ss_out[pj] = -acc;
                    P.z_out[pj] = zz;
                } else {
                for (int a = 0; a < K1; ++a) r[a] = auxs[a * KB_TN + tid] - f[a];
                double rmu = 0.0, muphi = 0.0;
                for (int a = 0; a < K1; ++a) {
                    double mu = 0.0;
                    for (int b = 0; b < K1; ++b) mu += Sinv[a * K1 + b] * r[b];
                    rmu += r[a] * mu;
                    muphi += mu * phi[a];
                }
                P.ss_out[pj] = P.vg.c0 - q + rmu;
                P.z_out[pj] = zc - muphi;
                }
            }
        }
        __syncthreads();      // qred / auxs / scratch are re-used by the next tile
    }
}

// Per-point finalize (deterministic reduction over row blocks + the (K+1)x(K+1) drift solve).
template <int DIM>
__global__ void __launch_bounds__(256) finalize_kernel(const __grid_constant__ FinalizeParams P) {
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.m) return;
    double q = 0.0;
    for (int I = 0; I < P.nrb; ++I) q += P.partial[(size_t)I * P.mpad + p];
    const int K = P.n_rl + P.n_hd, K1 = K + 1;
    double r[KB200_MAX_DRIFT + 1];
    double f[KB200_MAX_DRIFT + 1];
    if (P.n_rl > 0) {
        double x, y, z;
        kb_load_point<DIM>(P.ps, P.an, p, x, y, z);
        f[0] = (x - P.ds.shift[0]) * P.ds.scale[0];
        f[1] = (y - P.ds.shift[1]) * P.ds.scale[1];
        if (DIM == 3) f[2] = (z - P.ds.shift[2]) * P.ds.scale[2];
    }
    for (int c = 0; c < P.n_hd; ++c) {
        double v = P.drift_pts[(size_t)c * P.drift_stride + P.drift_first + p];
        f[P.n_rl + c] = (v - P.ds.shift[P.n_rl + c]) * P.ds.scale[P.n_rl + c];
    }
    f[K] = 1.0;
    for (int a = 0; a < K1; ++a) r[a] = P.auxout[(size_t)a * P.mpad + p] - f[a];
    const double zc = P.auxout[(size_t)K1 * P.mpad + p];
    const double* Sinv = P.consts;
    const double* phi = P.consts + K1 * K1;
    double rmu = 0.0, muphi = 0.0;
    for (int a = 0; a < K1; ++a) {
        double mu = 0.0;
        for (int b = 0; b < K1; ++b) mu += Sinv[a * K1 + b] * r[b];
        rmu += r[a] * mu;
        muphi += mu * phi[a];
    }
    P.ss_out[p] = P.vg.c0 - q + rmu;
    P.z_out[p] = zc - muphi;
}

size_t kbk_solve_smem(int dtype) {
    (void)dtype;
    return (size_t)SV_STAGES * (KB_BM * KB_BK + KB_BK * KB_TN) * sizeof(double) + 8 * KB_TN * sizeof(double) +
           SV_STAGES * sizeof(uint64_t) + 64;
}
// A/B switch for profiling only: KB200_SOLVE_V1=1 selects the non-specialised kernel.
static bool use_v1() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("KB200_SOLVE_V1"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}
static size_t solve_smem_pt() {
    return (size_t)PT_STAGES * (KB_BM * KB_BK + KB_BK * KB_TN) * sizeof(double) + 8 * KB_TN * sizeof(double) +
           KB_MAXAUX * KB_TN * sizeof(double) + 2 * PT_STAGES * sizeof(uint64_t) + 64;
}
bool kbk_solve_use_v1() { return use_v1(); }
size_t kbk_solve_pt_scratch_doubles(int n, int grid) {
    return (size_t)grid * ((n + KB_BK - 1) / KB_BK) * (KB_BK * KB_TN);
}


template <int DIM, int MODEL>
static cudaError_t solve_set_attr() {
    KB_CUDA_OK(cudaFuncSetAttribute(solve_kernel_pt<DIM, MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)solve_smem_pt()));
    return cudaFuncSetAttribute(solve_kernel_f64<DIM, MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)kbk_solve_smem(KB200_F64));
}

cudaError_t kbk_solve_init() {
#define KB_ATTR(M) KB_CUDA_OK((solve_set_attr<2, M>())); KB_CUDA_OK((solve_set_attr<3, M>())); KB_CUDA_OK((solve_set_attr<KB_GEO, M>()));
    KB_ATTR(KB200_VG_LINEAR) KB_ATTR(KB200_VG_POWER) KB_ATTR(KB200_VG_GAUSSIAN)
    KB_ATTR(KB200_VG_EXPONENTIAL) KB_ATTR(KB200_VG_SPHERICAL) KB_ATTR(KB200_VG_HOLE_EFFECT) KB_ATTR(KB200_VG_TABLE)
#undef KB_ATTR
    return cudaSuccess;
}

template <int DIM>
static cudaError_t solve_dim(int dtype, const SolveParams& p, cudaStream_t st) {
    if (dtype != KB200_F64) return cudaErrorNotSupported;
    dim3 grid((unsigned)(p.mpad / KB_TN), (unsigned)p.nrb);
    size_t sm = kbk_solve_smem(dtype);
    switch (p.vg.model) {
#define KB_CASE(M) case M: solve_kernel_f64<DIM, M><<<grid, SV_THREADS, sm, st>>>(p); break;
        KB_CASE(KB200_VG_LINEAR) KB_CASE(KB200_VG_POWER) KB_CASE(KB200_VG_GAUSSIAN)
        KB_CASE(KB200_VG_EXPONENTIAL) KB_CASE(KB200_VG_SPHERICAL) KB_CASE(KB200_VG_HOLE_EFFECT) KB_CASE(KB200_VG_TABLE)
#undef KB_CASE
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t kbk_solve(int dim, int dtype, const SolveParams& p, cudaStream_t st) {
    if (dim == KB_GEO) return solve_dim<KB_GEO>(dtype, p, st);
    return dim == 2 ? solve_dim<2>(dtype, p, st) : solve_dim<3>(dtype, p, st);
}

template <int DIM>
static cudaError_t solve_pt_dim(const SolvePtParams& p, int grid, cudaStream_t st) {
    size_t sm = solve_smem_pt();
    switch (p.vg.model) {
#define KB_CASE(M) case M: solve_kernel_pt<DIM, M><<<grid, PT_THREADS, sm, st>>>(p); break;
        KB_CASE(KB200_VG_LINEAR) KB_CASE(KB200_VG_POWER) KB_CASE(KB200_VG_GAUSSIAN)
        KB_CASE(KB200_VG_EXPONENTIAL) KB_CASE(KB200_VG_SPHERICAL) KB_CASE(KB200_VG_HOLE_EFFECT) KB_CASE(KB200_VG_TABLE)
#undef KB_CASE
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t kbk_solve_pt(int dim, const SolvePtParams& p, int grid, cudaStream_t st) {
    if (dim == KB_GEO) return solve_pt_dim<KB_GEO>(p, grid, st);
    return dim == 2 ? solve_pt_dim<2>(p, grid, st) : solve_pt_dim<3>(p, grid, st);
}

cudaError_t kbk_finalize(const FinalizeParams& p, cudaStream_t st) {
    unsigned g = (unsigned)((p.m + 255) / 256);
    if (p.dim == 2) finalize_kernel<2><<<g, 256, 0, st>>>(p);
    else if (p.dim == 3) finalize_kernel<3><<<g, 256, 0, st>>>(p);
    else finalize_kernel<KB_GEO><<<g, 256, 0, st>>>(p);
    return cudaGetLastError();
}
