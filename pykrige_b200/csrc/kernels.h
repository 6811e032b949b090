// kernels.h — internal launcher interface between api.cu and the kernel files.
#pragma once
#include "common.cuh"

struct DriftScale {            // f' = (f - shift) * scale  (change of drift basis)
    double shift[KB200_MAX_DRIFT + 1];
    double scale[KB200_MAX_DRIFT + 1];
};

// Tile stream layout of the packed inverse factor: row block I owns ktiles[I]
// consecutive tiles of KB_BM x KB_BK values starting at tile index tile_off[I].
struct PackMap {
    int nrb;
    int ktiles[KB_MAXRB];
    long long tile_off[KB_MAXRB];
};

struct SolveParams {
    VgParams vg;
    Aniso an;
    PointSource ps;
    int n, n_pad, na, nrb;
    const double* ax; const double* ay; const double* az;   // adjusted data coordinates
    const void* tiles;
    PackMap pm;
    long long m;          // points in this launch
    long long mpad;       // m rounded up to KB_TN (row stride of partial / auxout)
    double* partial;      // [nrb][mpad]   per row block sum of squares
    double* auxout;       // [na][mpad]    dual-row dot products
};

// K3 v3 (persistent point-tile kernel): solve + finalize in one launch
struct SolvePtParams {
    VgParams vg;
    Aniso an;
    PointSource ps;
    int n, na, nrb, n_rl, n_hd;
    const double* ax; const double* ay; const double* az;
    const void* tiles;
    PackMap pm;
    DriftScale ds;
    const double* consts;
    const double* drift_pts; long long drift_stride, drift_first;
    long long m;
    int gform;                // 1: tiles hold the symmetric inverse (quadratic form q = c^T G c), 0: W = chol(C)^-1
    const double* rowscale;   // int8-slice path only: 2^(ew_r - 12) per packed row
    double* scratch;          // [grid][ceil(n/16)][16*64]  RHS column blocks in fragment order
    double* z_out; double* ss_out;
};

struct FinalizeParams {
    VgParams vg;
    Aniso an;
    PointSource ps;
    int dim, n_rl, n_hd, nrb;
    DriftScale ds;
    const double* consts;      // Sinv (K1*K1), phi (K1)
    const double* drift_pts;   // device, column-major [n_hd][m_total] or null
    long long drift_stride;    // column stride of drift_pts
    long long drift_first;     // index of point 0 of this launch within drift_pts columns
    long long m, mpad;
    const double* partial; const double* auxout;
    double* z_out; double* ss_out;   // already offset to this launch's first point
};

cudaError_t kbk_adjust_data(int dim, const Aniso& an, int n, const double* x, const double* y, const double* z,
                            double* ax, double* ay, double* az, cudaStream_t st);
cudaError_t kbk_assemble(int dim, const VgParams& vg, int n, int n_pad, int ld,
                         const double* ax, const double* ay, const double* az, double* C, cudaStream_t st);
cudaError_t kbk_cholesky(double* C, double* W, int ld, int n_pad, int* flag, double dtol, cudaStream_t st, int* launches);
cudaError_t kbk_trtri(const double* L, double* W, double* T1, int ld, int n_pad, cudaStream_t st, int* launches);
cudaError_t kbk_dual(const double* W, int ld, int n, int n_pad, int n_rl, int n_hd,
                     const double* ax, const double* ay, const double* az, const DriftScale& ds,
                     const double* hd, const double* values,
                     double* Fz, double* Hz, double* Uz, double* consts, int* flag, cudaStream_t st, int* launches);
cudaError_t kbk_pack(int dtype, const double* W, int ld, int n, int n_pad, int na, const double* Uz,
                     const PackMap& pm, void* out, cudaStream_t st);

cudaError_t kbk_general_inverse(double* C, int ld, int n, int n_pad, double* rowbuf, double* colbuf, int* piv,
                                int* flag, double ptol, cudaStream_t st, int* launches);
cudaError_t kbk_dual_gform(const double* G, int ld, int n, int n_pad, int n_rl, int n_hd,
                           const double* ax, const double* ay, const double* az, const DriftScale& ds,
                           const double* hd, const double* values,
                           double* Fz, double* Uz, double* consts, int* flag, cudaStream_t st, int* launches);
cudaError_t kbk_pack_gform(const double* G, int ld, int n, int n_pad, int na, const double* Uz,
                           const PackMap& pm, void* out, cudaStream_t st);
cudaError_t kbk_factor_init();
cudaError_t kbk_solve_init();   // opt-in shared memory attributes
size_t      kbk_solve_smem(int dtype);
cudaError_t kbk_solve(int dim, int dtype, const SolveParams& p, cudaStream_t st);
cudaError_t kbk_finalize(const FinalizeParams& p, cudaStream_t st);
cudaError_t kbk_solve_pt(int dim, const SolvePtParams& p, int grid, cudaStream_t st);
bool        kbk_solve_use_v1();
size_t      kbk_solve_pt_scratch_doubles(int n, int grid);

// fp32 path (solve_tf32.cu): tcgen05.mma kind::tf32, 3xTF32 split, TMEM accumulators
cudaError_t kbk_solve_tf32_init();
cudaError_t kbk_solve_tf32(int dim, const SolvePtParams& p, int grid, cudaStream_t st);
cudaError_t kbk_pack_tf32(const double* W, int ld, int n, int n_pad, int na, const double* Uz, const PackMap& pm,
                          void* out, cudaStream_t st);
size_t      kbk_solve_tf32_scratch_bytes(int n, int grid);
int         kbk_solve_tf32_tile_points();

// fp64-class path on the INT8 tensor cores (solve_i8.cu): error-free slicing + exact int32 accumulation
cudaError_t kbk_solve_i8_init();
cudaError_t kbk_solve_i8(int dim, const SolvePtParams& p, int grid, cudaStream_t st);
cudaError_t kbk_pack_i8(const double* W, int ld, int n, int n_pad, int na, const double* Uz,
                        int* rowexp, double* rowscale, const long long* tile_off_dev, void* out, cudaStream_t st);
int         kbk_i8_nrb(int n, int na);
int         kbk_i8_rows(int n, int na);
long long   kbk_i8_total_tiles(int n, int na, long long* tile_off);
size_t      kbk_i8_tile_bytes();
size_t      kbk_solve_i8_scratch_bytes(int n, int grid);
int         kbk_solve_i8_tile_points();

// moving window (knn.cu)
struct KnnParams {
    VgParams vg;
    Aniso an;
    PointSource ps;
    int dim, n, k;
    const double* ax; const double* ay; const double* az; const double* values;  // adjusted data, cell-sorted
    const int* sorig;          // original index of each sorted point (deterministic tie-break)
    // uniform cell grid over the adjusted data
    int gx, gy, gz;
    int r0;                    // start radius of the cell search (from the mean point density)
    double ox, oy, oz, inv_cell, cell;
    const int* cell_start;     // [ncells+1]
    long long m;
    double* z_out; double* ss_out;
    int* flag;                 // singular local system
};
cudaError_t kbk_knn_build(int dim, int n, const double* ax, const double* ay, const double* az, const double* values,
                          KnnParams& kp, double* sx, double* sy, double* sz, double* sv, int* sorig,
                          int* cell_of, int* cell_start, int* cursor, int ncells, cudaStream_t st, int* launches);
cudaError_t kbk_knn_solve(const KnnParams& p, int chol, cudaStream_t st);
size_t      kbk_knn_smem_per_warp(int k, int chol);

// variogram.cu: constructor-side kernels (experimental variogram binning, cross-validation residuals)
cudaError_t kbk_ev_init();
int         kbk_ev_grid(int n, int num_sms);
size_t      kbk_ev_smem(int nlags, int priv);
int         kbk_ev_priv_max_lags();
cudaError_t kbk_ev_minmax(int dim, int n, const double* x, const double* y, const double* z, int grid,
                          double* bmin, double* bmax, cudaStream_t st);
cudaError_t kbk_ev_bin(int dim, int n, const double* x, const double* y, const double* z, const double* v,
                       int nlags, const double* edges, double inv_dd, int grid, double* part, double* out,
                       cudaStream_t st);
cudaError_t kbk_statistics(int dim, int n, const double* ax, const double* ay, const double* az,
                           const double* L, int ld, const double* u, const double* zeta, int* dup,
                           double* delta, double* sigma, cudaStream_t st);

// pinv.cu: pseudo_inv=True (one-sided Jacobi SVD of the bordered kriging matrix)
cudaError_t kbk_build_fz(int n, int n_pad, int n_rl, int n_hd, const double* ax, const double* ay, const double* az,
                         const DriftScale& ds, const double* hd, const double* values, double* Fz, cudaStream_t st);
cudaError_t kbk_pinv_init();
int         kbk_pinv_max_nt();
size_t      kbk_pinv_workspace_doubles(int nt);
cudaError_t kbk_pinv(int n, int K1, int n_pad, double* C, int ldc, const double* Fz, const double* values,
                     double* Uz, double* consts, double* work, int* counter, cudaStream_t st,
                     int* launches, int* sweeps, int* rank);
