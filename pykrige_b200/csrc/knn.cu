// knn.cu — moving-window kriging (n_closest_points): exact kNN on a uniform cell grid +
// per-point local (k+1)x(k+1) solve.  (ok.py:722-758, 929-986; cok.pyx:98-193)
#include "common.cuh"
#include "kernels.h"

struct kb200_ctx;
extern "C" int kb200_set_problem_knn(kb200_handle, int, int64_t, const double*, const double*, const double*,
                                     const double*, const double*, const double*, int, const double*, int, int, double) {
    return KB200_EUNSUPPORTED;
}
extern "C" int kb200_execute_knn_points(kb200_handle, int, int64_t, const double*, const double*, const double*,
                                        double*, double*) { return KB200_EUNSUPPORTED; }
extern "C" int kb200_execute_knn_grid(kb200_handle, int, int64_t, int64_t, int64_t, const double*, const double*,
                                      const double*, int64_t, int64_t, double*, double*) { return KB200_EUNSUPPORTED; }
extern "C" int kb200_execute_knn_grid_dev(kb200_handle, int, int64_t, int64_t, int64_t, const double*, const double*,
                                          const double*, int64_t, int64_t, double*, double*) { return KB200_EUNSUPPORTED; }
