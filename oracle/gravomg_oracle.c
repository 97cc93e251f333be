/* gravomg_oracle.c — CPU restatement of Gravo MG's solve path. TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, not the product: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it. Nothing under gravo_mg_b200/
 * links, imports or calls it.
 *
 * PARITY UNPINNED: the reference repository (rubenwiersma/gravo_mg @ d0126770) holds no tests,
 * golden vectors or fixtures for this path, and it cannot be built here (its CMake fetches
 * libigl v2.3.0 and Eigen 3.3.x from the network; neither is vendored nor installed). This
 * restatement follows the reference source line by line where the source is visible and the
 * published algorithm where the arithmetic lives in Eigen:
 *
 *   orc_gauss_seidel     gravomg/src/multigrid_solver.cpp:1194-1226  (hand-written, fully visible)
 *   orc_vcycle           multigrid_solver.cpp:1059-1088
 *   orc_residual_check   multigrid_solver.cpp:1228-1277
 *   orc_setup/orc_solve  multigrid_solver.cpp:1279-1285, 1367-1449; x0 = rhs is the caller's
 *                        job (gravomg_bindings/src/cpp/core.cpp:69)
 *   csc_times_dense      Eigen 3.3 SparseDenseProduct, ColMajor lhs: for each column j, for each
 *                        stored (i,j): res(i) += A(i,j)*x(j)            (call sites :1066, :1082)
 *   csct_times_dense     Eigen 3.3 SparseDenseProduct, Transpose<ColMajor>: one dot product per
 *                        output row over the column of U                 (call site :1069)
 *   csc_spgemm           Eigen 3.3 conservative_sparse_sparse_product: column j of the result
 *                        accumulates lhs(:,k)*rhs(k,j) over the stored k of rhs(:,j) in order
 *                        (call sites :1389-1391, evaluated left to right: (U^T A) U)
 *   ldl_*                Eigen::SimplicialLDLT follows T. Davis' LDL (ACM TOMS 31(4), 2005):
 *                        elimination tree + up-looking numeric factorisation, restated here.
 *                        Eigen orders with AMD; this file uses reverse Cuthill-McKee. The
 *                        ordering only changes rounding (the exact-arithmetic solution is the
 *                        same); the parity tests therefore compare coarse solves to 1e-9.
 *
 *   *_diff               NOT in the reference: the device path's cancellation-free row product on the
 *                        finest level, sum_{j != i} A_ij (x_j - x_i) + s_i x_i with s_i the row sum of
 *                        A accumulated with an error-free TwoSum chain (orc_rowsum). Algebraically the
 *                        same A x; on Poisson systems tau M + S the iterate carries a constant of size
 *                        1/(tau sqrt N) whose products cancel in the plain form and leave a rounding
 *                        floor of ~1e-6 in the relative residual at >= 4 M vertices. Only the Jacobi
 *                        variant (the device's cycle) uses it; the Gauss-Seidel reference path does not.
 *
 * Conventions: sparse matrices are CSC with int32 indices (what the pybind11 Eigen caster
 * hands the reference); dense blocks are column-major n x K (Eigen::MatrixXd).
 * Compile with -ffp-contract=off so a*b+c is never fused (the device kernels are compiled
 * with -fmad=false for the same reason).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define ORC_MAX_LEVELS 16
#define ORC_MAX_SWEEPS 16

typedef struct {
    int rows, cols;
    int* colptr;
    int* rowidx;
    double* vals;
} csc_t;

typedef struct {
    int n;
    int* perm;   /* new -> old */
    int* Lp;
    int* Li;
    double* Lx;
    double* D;
    int* parent;
} ldl_t;

typedef struct orc_solver {
    int n;
    double* mass;                 /* lumped mass diagonal (constructor argument) */
    int n_levels;                 /* number of prolongations U[k] */
    csc_t U[ORC_MAX_LEVELS];
    csc_t Ut[ORC_MAX_LEVELS];     /* explicit transposes, CSC (only used by the Galerkin product) */
    csc_t Abar[ORC_MAX_LEVELS + 1]; /* Abar[0] unused (level 0 is the caller's LHS) */
    ldl_t coarse;
    int have_setup;
    /* parameters the binding sets (core.cpp:52-57) */
    int pre_iters, post_iters, max_iter, criterion, smoother; /* smoother 0 = GS (reference), 1 = damped Jacobi */
    int cycle_type; /* 0 V, 1 F, 2 W (core.cpp:52; multigrid_solver.cpp:1420-1439) */
    int row_product; /* Jacobi variant only: 1 = cancellation-free row product on level 0 (the device default) */
    double tol, omega;
    double w_pre[ORC_MAX_LEVELS][ORC_MAX_SWEEPS], w_post[ORC_MAX_LEVELS][ORC_MAX_SWEEPS]; /* Jacobi damping per level and sweep */
    /* outputs */
    double t_reduction_ms, t_coarse_ms, t_cycles_ms, t_total_ms;
    int iterations;
    double residue;
} orc_solver;

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static void csc_free(csc_t* m) {
    free(m->colptr), free(m->rowidx), free(m->vals);
    memset(m, 0, sizeof *m);
}

static csc_t csc_copy(int rows, int cols, const int* cp, const int* ri, const double* v) {
    csc_t m;
    m.rows = rows, m.cols = cols;
    int nnz = cp[cols];
    m.colptr = (int*)malloc(sizeof(int) * (size_t)(cols + 1));
    m.rowidx = (int*)malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
    m.vals = (double*)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    memcpy(m.colptr, cp, sizeof(int) * (size_t)(cols + 1));
    memcpy(m.rowidx, ri, sizeof(int) * (size_t)nnz);
    memcpy(m.vals, v, sizeof(double) * (size_t)nnz);
    return m;
}

/* Transpose; row indices inside every output column come out ascending. */
static csc_t csc_transpose(const csc_t* a) {
    csc_t t;
    t.rows = a->cols, t.cols = a->rows;
    int nnz = a->colptr[a->cols];
    t.colptr = (int*)calloc((size_t)t.cols + 1, sizeof(int));
    t.rowidx = (int*)malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
    t.vals = (double*)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    for (int p = 0; p < nnz; ++p) t.colptr[a->rowidx[p] + 1]++;
    for (int j = 0; j < t.cols; ++j) t.colptr[j + 1] += t.colptr[j];
    int* cur = (int*)malloc(sizeof(int) * (size_t)(t.cols > 0 ? t.cols : 1));
    memcpy(cur, t.colptr, sizeof(int) * (size_t)t.cols);
    for (int j = 0; j < a->cols; ++j)
        for (int p = a->colptr[j]; p < a->colptr[j + 1]; ++p) {
            int q = cur[a->rowidx[p]]++;
            t.rowidx[q] = j;
            t.vals[q] = a->vals[p];
        }
    free(cur);
    return t;
}

static int cmp_int(const void* a, const void* b) { return (*(const int*)a > *(const int*)b) - (*(const int*)a < *(const int*)b); }

/* C = A * B, Eigen's conservative sparse product restated: per result column j a dense
 * accumulator indexed by row; the stored k of B(:,j) are visited in order and for each the
 * stored rows of A(:,k); result rows sorted ascending. */
static csc_t csc_spgemm(const csc_t* A, const csc_t* B) {
    csc_t C;
    C.rows = A->rows, C.cols = B->cols;
    C.colptr = (int*)calloc((size_t)C.cols + 1, sizeof(int));
    size_t cap = (size_t)(A->colptr[A->cols] + B->colptr[B->cols]) + 16;
    C.rowidx = (int*)malloc(sizeof(int) * cap);
    C.vals = (double*)malloc(sizeof(double) * cap);
    double* acc = (double*)calloc((size_t)(A->rows > 0 ? A->rows : 1), sizeof(double));
    char* mask = (char*)calloc((size_t)(A->rows > 0 ? A->rows : 1), 1);
    int* list = (int*)malloc(sizeof(int) * (size_t)(A->rows > 0 ? A->rows : 1));
    size_t nnz = 0;
    for (int j = 0; j < B->cols; ++j) {
        int count = 0;
        for (int q = B->colptr[j]; q < B->colptr[j + 1]; ++q) {
            const int k = B->rowidx[q];
            const double y = B->vals[q];
            for (int p = A->colptr[k]; p < A->colptr[k + 1]; ++p) {
                const int i = A->rowidx[p];
                const double x = A->vals[p];
                if (!mask[i]) {
                    mask[i] = 1;
                    acc[i] = x * y;
                    list[count++] = i;
                } else {
                    acc[i] += x * y;
                }
            }
        }
        qsort(list, (size_t)count, sizeof(int), cmp_int);
        if (nnz + (size_t)count > cap) {
            cap = (nnz + (size_t)count) * 2;
            C.rowidx = (int*)realloc(C.rowidx, sizeof(int) * cap);
            C.vals = (double*)realloc(C.vals, sizeof(double) * cap);
        }
        for (int t = 0; t < count; ++t) {
            const int i = list[t];
            C.rowidx[nnz] = i;
            C.vals[nnz] = acc[i];
            ++nnz;
            mask[i] = 0;
        }
        C.colptr[j + 1] = (int)nnz;
    }
    free(acc), free(mask), free(list);
    return C;
}

/* res += A * x, one dense column (Eigen ColMajor sparse * dense: scatter / axpy form). */
static void csc_times_dense_add(const csc_t* A, const double* x, double* res) {
    for (int j = 0; j < A->cols; ++j) {
        const double xj = x[j];
        for (int p = A->colptr[j]; p < A->colptr[j + 1]; ++p) res[A->rowidx[p]] += A->vals[p] * xj;
    }
}

/* res = A^T * x, one dense column (row-major view: one dot product per output entry). */
static void csct_times_dense(const csc_t* A, const double* x, double* res) {
    for (int j = 0; j < A->cols; ++j) {
        double s = 0.0;
        for (int p = A->colptr[j]; p < A->colptr[j + 1]; ++p) s += A->vals[p] * x[A->rowidx[p]];
        res[j] = s;
    }
}

/* LHS.coeffRef(k,k) on a compressed matrix: binary search of row k in column k (0 when absent;
 * the reference would insert it). */
static double csc_diag(const csc_t* A, int k) {
    int lo = A->colptr[k], hi = A->colptr[k + 1] - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        if (A->rowidx[mid] == k) return A->vals[mid];
        if (A->rowidx[mid] < k) lo = mid + 1; else hi = mid - 1;
    }
    return 0.0;
}

/* ------------------------------------------------------------------ smoothers */
/* multigrid_solver.cpp:1194-1226: lexicographic Gauss-Seidel, in place, column k of the CSC
 * matrix used as row k (valid for symmetric systems), columns of x one at a time. */
void orc_gauss_seidel(int n, const int* cp, const int* ri, const double* v, const double* rhs, double* x, int K,
                      int iters) {
    csc_t A = {n, n, (int*)cp, (int*)ri, (double*)v};
    for (int it = 0; it < iters; ++it)
        for (int c = 0; c < K; ++c) {
            double* xc = x + (size_t)c * n;
            const double* bc = rhs + (size_t)c * n;
            for (int k = 0; k < n; ++k) {
                double sum = 0.0;
                for (int p = cp[k]; p < cp[k + 1]; ++p)
                    if (ri[p] != k) sum += v[p] * xc[ri[p]];
                xc[k] = (bc[k] - sum) / csc_diag(&A, k);
            }
        }
}

/* Damped Jacobi, the sweep the device path substitutes: x' = x + (omega / A_kk) * (b - (A x)_k),
 * (A x)_k summed over the stored entries of column k (= row k, symmetric A) in stored order,
 * diagonal included. `tmp` is n x K scratch. Same association as the CUDA epilogue.
 * omegas[it] is the damping of sweep `it` (a constant list is plain damped Jacobi; Chebyshev
 * roots give the polynomial smoother the device uses by default). */
void orc_jacobi(int n, const int* cp, const int* ri, const double* v, const double* rhs, double* x, double* tmp, int K,
                int iters, const double* omegas) {
    for (int it = 0; it < iters; ++it) {
        const double omega = omegas[it];
        for (int c = 0; c < K; ++c) {
            const double* xc = x + (size_t)c * n;
            const double* bc = rhs + (size_t)c * n;
            double* tc = tmp + (size_t)c * n;
            for (int k = 0; k < n; ++k) {
                double acc = 0.0, d = 0.0;
                for (int p = cp[k]; p < cp[k + 1]; ++p) {
                    acc += v[p] * xc[ri[p]];
                    if (ri[p] == k) d += v[p];
                }
                const double s = omega * (1.0 / d);
                tc[k] = xc[k] + s * (bc[k] - acc);
            }
        }
        memcpy(x, tmp, sizeof(double) * (size_t)n * K);
    }
}

/* ------------------------------------------------------------------ cancellation-free row product (device path) */
/* s_k = sum of the stored entries of column k (= row k), accumulated in stored order with Knuth's
 * TwoSum: `s` is the running floating-point sum, `e` collects the rounding error of every addition,
 * the result is s + e. Exact to ~eps^2 sum|A_kj|: the row sums of tau M + S are ~1e-12 next to entries ~1. */
void orc_rowsum(int n, const int* cp, const double* v, double* out) {
    for (int k = 0; k < n; ++k) {
        double s = 0.0, e = 0.0;
        for (int p = cp[k]; p < cp[k + 1]; ++p) {
            const double a = v[p];
            const double t = s + a;
            const double bp = t - s;
            const double err = (s - (t - bp)) + (a - bp);
            s = t;
            e += err;
        }
        out[k] = s + e;
    }
}

/* (A x)_k = sum_p term_p in stored order: off-diagonal entries contribute A_kj (x_j - x_k), the first
 * stored diagonal entry contributes s_k x_k, further (duplicate) diagonal entries nothing. */
static double row_product_diff(const int* cp, const int* ri, const double* v, const double* rowsum, const double* xc, int k) {
    double acc = 0.0;
    const double xk = xc[k];
    int seen = 0;
    for (int p = cp[k]; p < cp[k + 1]; ++p) {
        if (ri[p] == k) {
            acc += (seen ? 0.0 : rowsum[k]) * xk;
            seen = 1;
        } else {
            acc += v[p] * (xc[ri[p]] - xk);
        }
    }
    return acc;
}

void orc_jacobi_diff(int n, const int* cp, const int* ri, const double* v, const double* rhs, double* x, double* tmp, int K,
                     int iters, const double* omegas) {
    double* rowsum = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    orc_rowsum(n, cp, v, rowsum);
    for (int it = 0; it < iters; ++it) {
        const double omega = omegas[it];
        for (int c = 0; c < K; ++c) {
            const double* xc = x + (size_t)c * n;
            const double* bc = rhs + (size_t)c * n;
            double* tc = tmp + (size_t)c * n;
            for (int k = 0; k < n; ++k) {
                double d = 0.0;
                for (int p = cp[k]; p < cp[k + 1]; ++p)
                    if (ri[p] == k) d += v[p];
                const double acc = row_product_diff(cp, ri, v, rowsum, xc, k);
                const double s = omega * (1.0 / d);
                tc[k] = xc[k] + s * (bc[k] - acc);
            }
        }
        memcpy(x, tmp, sizeof(double) * (size_t)n * K);
    }
    free(rowsum);
}

void orc_residual_diff(int n, const int* cp, const int* ri, const double* v, const double* b, const double* x, double* res,
                       int K) {
    double* rowsum = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    orc_rowsum(n, cp, v, rowsum);
    for (int c = 0; c < K; ++c)
        for (int i = 0; i < n; ++i)
            res[(size_t)c * n + i] = b[(size_t)c * n + i] - row_product_diff(cp, ri, v, rowsum, x + (size_t)c * n, i);
    free(rowsum);
}

/* ------------------------------------------------------------------ single operators */
/* res = b - A*x (multigrid_solver.cpp:1066). */
void orc_residual(int n, const int* cp, const int* ri, const double* v, const double* b, const double* x, double* res,
                  int K) {
    csc_t A = {n, n, (int*)cp, (int*)ri, (double*)v};
    double* ax = (double*)malloc(sizeof(double) * (size_t)n);
    for (int c = 0; c < K; ++c) {
        memset(ax, 0, sizeof(double) * (size_t)n);
        csc_times_dense_add(&A, x + (size_t)c * n, ax);
        for (int i = 0; i < n; ++i) res[(size_t)c * n + i] = b[(size_t)c * n + i] - ax[i];
    }
    free(ax);
}

/* out = U^T * r (multigrid_solver.cpp:1069). U is rows x cols CSC. */
void orc_restrict(int rows, int cols, const int* cp, const int* ri, const double* v, const double* r, double* out,
                  int K) {
    csc_t U = {rows, cols, (int*)cp, (int*)ri, (double*)v};
    for (int c = 0; c < K; ++c) csct_times_dense(&U, r + (size_t)c * rows, out + (size_t)c * cols);
}

/* x = x + U * eps (multigrid_solver.cpp:1082). */
void orc_prolong_add(int rows, int cols, const int* cp, const int* ri, const double* v, const double* eps, double* x,
                     int K) {
    csc_t U = {rows, cols, (int*)cp, (int*)ri, (double*)v};
    double* ue = (double*)malloc(sizeof(double) * (size_t)rows);
    for (int c = 0; c < K; ++c) {
        memset(ue, 0, sizeof(double) * (size_t)rows);
        csc_times_dense_add(&U, eps + (size_t)c * cols, ue);
        for (int i = 0; i < rows; ++i) x[(size_t)c * rows + i] = x[(size_t)c * rows + i] + ue[i];
    }
    free(ue);
}

/* residualCheck (multigrid_solver.cpp:1228-1277). mass may be NULL for types 0 and 3.
 * Minv = igl::invert_diag(M): 1/m where m != 0 (multigrid_solver.cpp:19). */
double orc_residual_check(int n, const int* cp, const int* ri, const double* v, const double* b, const double* x, int K,
                          int type, const double* mass) {
    csc_t A = {n, n, (int*)cp, (int*)ri, (double*)v};
    double* r = (double*)malloc(sizeof(double) * (size_t)n);
    double best = 0.0, frob = 0.0;
    for (int c = 0; c < K; ++c) {
        const double* bc = b + (size_t)c * n;
        memset(r, 0, sizeof(double) * (size_t)n);
        csc_times_dense_add(&A, x + (size_t)c * n, r);
        double n1 = 0.0, n2 = 0.0;
        for (int i = 0; i < n; ++i) {
            const double ri_ = r[i] - bc[i];
            double w = 1.0;
            if (type == 2) w = mass[i];
            if (type == 1) w = mass[i] != 0.0 ? 1.0 / mass[i] : 0.0;
            n1 += ri_ * (w * ri_);
            n2 += bc[i] * (w * bc[i]);
        }
        frob += n1;
        const double rel = sqrt(n1 / n2);
        if (c == 0 || rel > best) best = rel;
    }
    free(r);
    return type == 3 ? sqrt(frob) : best;
}

/* residualCheck with the cancellation-free row product (same norms, same association). */
double orc_residual_check_diff(int n, const int* cp, const int* ri, const double* v, const double* b, const double* x, int K,
                               int type, const double* mass) {
    double* rowsum = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    orc_rowsum(n, cp, v, rowsum);
    double best = 0.0, frob = 0.0;
    for (int c = 0; c < K; ++c) {
        const double* bc = b + (size_t)c * n;
        double n1 = 0.0, n2 = 0.0;
        for (int i = 0; i < n; ++i) {
            const double ri_ = row_product_diff(cp, ri, v, rowsum, x + (size_t)c * n, i) - bc[i];
            double w = 1.0;
            if (type == 2) w = mass[i];
            if (type == 1) w = mass[i] != 0.0 ? 1.0 / mass[i] : 0.0;
            n1 += ri_ * (w * ri_);
            n2 += bc[i] * (w * bc[i]);
        }
        frob += n1;
        const double rel = sqrt(n1 / n2);
        if (c == 0 || rel > best) best = rel;
    }
    free(rowsum);
    return type == 3 ? sqrt(frob) : best;
}

/* ------------------------------------------------------------------ sparse LDL^T (Davis' LDL) */
static void ldl_free(ldl_t* f) {
    free(f->perm), free(f->Lp), free(f->Li), free(f->Lx), free(f->D), free(f->parent);
    memset(f, 0, sizeof *f);
}

typedef struct { int deg, v; } degv_t;
static int cmp_degv(const void* a, const void* b) {
    const degv_t* x = (const degv_t*)a; const degv_t* y = (const degv_t*)b;
    if (x->deg != y->deg) return (x->deg > y->deg) - (x->deg < y->deg);
    return (x->v > y->v) - (x->v < y->v);
}

/* Reverse Cuthill-McKee on the pattern of A (assumed structurally symmetric). */
static void rcm_order(const csc_t* A, int* perm) {
    const int n = A->cols;
    char* seen = (char*)calloc((size_t)(n > 0 ? n : 1), 1);
    int* order = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    degv_t* buf = (degv_t*)malloc(sizeof(degv_t) * (size_t)(n > 0 ? n : 1));
    int filled = 0;
    while (filled < n) {
        int start = -1, best = 0x7fffffff;
        for (int i = 0; i < n; ++i)
            if (!seen[i] && A->colptr[i + 1] - A->colptr[i] < best) best = A->colptr[i + 1] - A->colptr[i], start = i;
        seen[start] = 1;
        order[filled++] = start;
        for (int head = filled - 1; head < filled; ++head) {
            const int u = order[head];
            int cnt = 0;
            for (int p = A->colptr[u]; p < A->colptr[u + 1]; ++p) {
                const int w = A->rowidx[p];
                if (!seen[w]) {
                    seen[w] = 1;
                    buf[cnt].deg = A->colptr[w + 1] - A->colptr[w];
                    buf[cnt].v = w;
                    ++cnt;
                }
            }
            qsort(buf, (size_t)cnt, sizeof(degv_t), cmp_degv);
            for (int t = 0; t < cnt; ++t) order[filled++] = buf[t].v;
        }
    }
    for (int i = 0; i < n; ++i) perm[i] = order[n - 1 - i];
    free(seen), free(order), free(buf);
}

/* Factor P A P^T = L D L^T. Returns 0 on success, k+1 when D(k) is zero. */
static int ldl_factor(const csc_t* A, ldl_t* f) {
    const int n = A->cols;
    memset(f, 0, sizeof *f);
    f->n = n;
    f->perm = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    rcm_order(A, f->perm);
    int* pinv = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) pinv[f->perm[i]] = i;
    /* C = upper part of P A P^T, CSC, columns in new numbering */
    int* Cp = (int*)calloc((size_t)n + 1, sizeof(int));
    for (int j = 0; j < n; ++j)
        for (int p = A->colptr[j]; p < A->colptr[j + 1]; ++p) {
            const int i2 = pinv[A->rowidx[p]], j2 = pinv[j];
            if (i2 <= j2) Cp[j2 + 1]++;
        }
    for (int j = 0; j < n; ++j) Cp[j + 1] += Cp[j];
    const int cnz = Cp[n];
    int* Ci = (int*)malloc(sizeof(int) * (size_t)(cnz > 0 ? cnz : 1));
    double* Cx = (double*)malloc(sizeof(double) * (size_t)(cnz > 0 ? cnz : 1));
    int* cur = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    memcpy(cur, Cp, sizeof(int) * (size_t)n);
    for (int j = 0; j < n; ++j)
        for (int p = A->colptr[j]; p < A->colptr[j + 1]; ++p) {
            const int i2 = pinv[A->rowidx[p]], j2 = pinv[j];
            if (i2 <= j2) {
                const int q = cur[j2]++;
                Ci[q] = i2;
                Cx[q] = A->vals[p];
            }
        }
    /* symbolic: elimination tree and column counts */
    int* parent = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int* lnz = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int* flag = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int k = 0; k < n; ++k) {
        parent[k] = -1, flag[k] = k, lnz[k] = 0;
        for (int p = Cp[k]; p < Cp[k + 1]; ++p) {
            int i = Ci[p];
            if (i < k)
                for (; flag[i] != k; i = parent[i]) {
                    if (parent[i] == -1) parent[i] = k;
                    lnz[i]++;
                    flag[i] = k;
                }
        }
    }
    f->Lp = (int*)malloc(sizeof(int) * ((size_t)n + 1));
    f->Lp[0] = 0;
    for (int k = 0; k < n; ++k) f->Lp[k + 1] = f->Lp[k] + lnz[k];
    const int lnnz = f->Lp[n];
    f->Li = (int*)malloc(sizeof(int) * (size_t)(lnnz > 0 ? lnnz : 1));
    f->Lx = (double*)malloc(sizeof(double) * (size_t)(lnnz > 0 ? lnnz : 1));
    f->D = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    f->parent = parent;
    /* numeric: up-looking, one sparse triangular solve per row of L */
    double* Y = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    int* pattern = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int status = 0;
    for (int k = 0; k < n && !status; ++k) {
        int top = n;
        flag[k] = k;
        lnz[k] = 0;
        for (int p = Cp[k]; p < Cp[k + 1]; ++p) {
            int i = Ci[p];
            if (i <= k) {
                Y[i] += Cx[p];
                int len = 0;
                for (; flag[i] != k; i = parent[i]) {
                    pattern[len++] = i;
                    flag[i] = k;
                }
                while (len > 0) pattern[--top] = pattern[--len];
            }
        }
        f->D[k] = Y[k];
        Y[k] = 0.0;
        for (; top < n; ++top) {
            const int i = pattern[top];
            const double yi = Y[i];
            Y[i] = 0.0;
            int p;
            for (p = f->Lp[i]; p < f->Lp[i] + lnz[i]; ++p) Y[f->Li[p]] -= f->Lx[p] * yi;
            const double lki = yi / f->D[i];
            f->D[k] -= lki * yi;
            f->Li[p] = k;
            f->Lx[p] = lki;
            lnz[i]++;
        }
        if (f->D[k] == 0.0) status = k + 1;
    }
    free(pinv), free(Cp), free(Ci), free(Cx), free(cur), free(lnz), free(flag), free(Y), free(pattern);
    return status;
}

/* x = A^-1 b for one column: permute, L solve, D^-1, L^T solve, permute back. */
static void ldl_solve(const ldl_t* f, const double* b, double* x, double* work) {
    const int n = f->n;
    for (int i = 0; i < n; ++i) work[i] = b[f->perm[i]];
    for (int j = 0; j < n; ++j)
        for (int p = f->Lp[j]; p < f->Lp[j + 1]; ++p) work[f->Li[p]] -= f->Lx[p] * work[j];
    for (int j = 0; j < n; ++j) work[j] /= f->D[j];
    for (int j = n - 1; j >= 0; --j)
        for (int p = f->Lp[j]; p < f->Lp[j + 1]; ++p) work[j] -= f->Lx[p] * work[f->Li[p]];
    for (int i = 0; i < n; ++i) x[f->perm[i]] = work[i];
}

/* ------------------------------------------------------------------ solver object */
orc_solver* orc_create(int n, const double* mass_diag) {
    orc_solver* s = (orc_solver*)calloc(1, sizeof(orc_solver));
    s->n = n;
    s->mass = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    memcpy(s->mass, mass_diag, sizeof(double) * (size_t)n);
    /* Python defaults, core.py:10 */
    s->pre_iters = 2, s->post_iters = 2, s->max_iter = 100, s->criterion = 2, s->tol = 1e-4;
    s->smoother = 0, s->omega = 2.0 / 3.0;
    for (int k = 0; k < ORC_MAX_LEVELS; ++k)
        for (int i = 0; i < ORC_MAX_SWEEPS; ++i) s->w_pre[k][i] = s->w_post[k][i] = s->omega;
    return s;
}

static void drop_setup(orc_solver* s) {
    for (int k = 0; k <= ORC_MAX_LEVELS; ++k)
        if (s->Abar[k].colptr) csc_free(&s->Abar[k]);
    if (s->coarse.perm) ldl_free(&s->coarse);
    s->have_setup = 0;
}

void orc_destroy(orc_solver* s) {
    if (!s) return;
    drop_setup(s);
    for (int k = 0; k < s->n_levels; ++k) csc_free(&s->U[k]), csc_free(&s->Ut[k]);
    free(s->mass);
    free(s);
}

int orc_add_prolongation(orc_solver* s, int rows, int cols, const int* cp, const int* ri, const double* v) {
    if (s->n_levels >= ORC_MAX_LEVELS) return 1;
    const int expect = s->n_levels == 0 ? s->n : s->U[s->n_levels - 1].cols;
    if (rows != expect) return 2;
    s->U[s->n_levels] = csc_copy(rows, cols, cp, ri, v);
    s->Ut[s->n_levels] = csc_transpose(&s->U[s->n_levels]);
    s->n_levels++;
    drop_setup(s);
    return 0;
}

void orc_set_params(orc_solver* s, int pre_iters, int post_iters, int max_iter, int criterion, double tol, int smoother,
                    double omega) {
    s->pre_iters = pre_iters, s->post_iters = post_iters, s->max_iter = max_iter, s->criterion = criterion;
    s->tol = tol, s->smoother = smoother, s->omega = omega;
    for (int k = 0; k < ORC_MAX_LEVELS; ++k)
        for (int i = 0; i < ORC_MAX_SWEEPS; ++i) s->w_pre[k][i] = s->w_post[k][i] = omega;
}

/* Per-sweep Jacobi damping of one level (pre- and post-smoothing sequences). */
int orc_set_weights(orc_solver* s, int level, int n_pre, const double* pre, int n_post, const double* post) {
    if (level < 0 || level >= ORC_MAX_LEVELS || n_pre > ORC_MAX_SWEEPS || n_post > ORC_MAX_SWEEPS) return 1;
    for (int i = 0; i < n_pre; ++i) s->w_pre[level][i] = pre[i];
    for (int i = 0; i < n_post; ++i) s->w_post[level][i] = post[i];
    return 0;
}

/* "reduction" + "coarsest_solve" of multigrid_solver.cpp:1387-1401. */
int orc_setup(orc_solver* s, const int* cp, const int* ri, const double* v) {
    drop_setup(s);
    csc_t A = {s->n, s->n, (int*)cp, (int*)ri, (double*)v};
    const double t0 = now_ms();
    const csc_t* cur = &A;
    for (int k = 0; k < s->n_levels; ++k) {
        csc_t T = csc_spgemm(&s->Ut[k], cur);       /* (U^T * A) ... */
        s->Abar[k + 1] = csc_spgemm(&T, &s->U[k]);  /* ... * U */
        csc_free(&T);
        cur = &s->Abar[k + 1];
    }
    const double t1 = now_ms();
    const int status = ldl_factor(cur, &s->coarse);
    const double t2 = now_ms();
    s->t_reduction_ms = t1 - t0;
    s->t_coarse_ms = t2 - t1;
    s->have_setup = 1;
    return status;
}

static void smooth(const orc_solver* s, const csc_t* A, const double* b, double* x, int K, int iters, const double* omegas, int level) {
    if (s->smoother == 0) {
        orc_gauss_seidel(A->cols, A->colptr, A->rowidx, A->vals, b, x, K, iters);
    } else {
        double* tmp = (double*)malloc(sizeof(double) * (size_t)A->cols * K);
        if (s->row_product && level == 0)
            orc_jacobi_diff(A->cols, A->colptr, A->rowidx, A->vals, b, x, tmp, K, iters, omegas);
        else
            orc_jacobi(A->cols, A->colptr, A->rowidx, A->vals, b, x, tmp, K, iters, omegas);
        free(tmp);
    }
}

/* type 0: multiGridVCycleGS (multigrid_solver.cpp:1059-1088); 1: multiGridFCycleGS (:1091-1140); 2: multiGridWCycleGS
 * (:1143-1192). The F- and W-cycles repeat residual / restriction / recursion / prolongation / post-smoothing; the
 * second recursion (a V-cycle for F, a W-cycle for W) starts from the eps of the first one (eps is not reset at :1126 /
 * :1178). Upstream tests `k == DoF.size() - 2` for the coarsest level there, which is off by one whenever the hierarchy
 * stopped at lowBound (DoF keeps the discarded level, :129,156-159) and then indexes past U; restated with the test of
 * the first recursion (`k == U.size() - 1`), the evident intent. */
static void cycle(const orc_solver* s, const csc_t* A, const double* b, double* x, int K, int k, int type) {
    const int n = A->cols;
    if (s->n_levels == 0) { /* undefined upstream (U[0] out of range); whole system to the direct solver */
        double* work = (double*)malloc(sizeof(double) * (size_t)n);
        for (int c = 0; c < K; ++c) ldl_solve(&s->coarse, b + (size_t)c * n, x + (size_t)c * n, work);
        free(work);
        return;
    }
    const csc_t* U = &s->U[k];
    const int nc = U->cols;
    smooth(s, A, b, x, K, s->pre_iters, s->w_pre[k], k);
    double* res = (double*)malloc(sizeof(double) * (size_t)n * K);
    double* rest = (double*)malloc(sizeof(double) * (size_t)nc * K);
    double* eps = (double*)calloc((size_t)nc * K, sizeof(double));
    for (int half = 0; half < (type == 0 ? 1 : 2); ++half) {
        if (s->smoother == 1 && s->row_product && k == 0)
            orc_residual_diff(n, A->colptr, A->rowidx, A->vals, b, x, res, K);
        else
            orc_residual(n, A->colptr, A->rowidx, A->vals, b, x, res, K);
        orc_restrict(n, nc, U->colptr, U->rowidx, U->vals, res, rest, K);
        if (k == s->n_levels - 1) {
            double* work = (double*)malloc(sizeof(double) * (size_t)nc);
            for (int c = 0; c < K; ++c) ldl_solve(&s->coarse, rest + (size_t)c * nc, eps + (size_t)c * nc, work);
            free(work);
        } else {
            /* first recursion: the same cycle type; second: V inside an F-cycle, W inside a W-cycle */
            cycle(s, &s->Abar[k + 1], rest, eps, K, k + 1, half == 0 ? type : (type == 1 ? 0 : 2));
        }
        orc_prolong_add(n, nc, U->colptr, U->rowidx, U->vals, eps, x, K);
        smooth(s, A, b, x, K, s->post_iters, s->w_post[k], k);
    }
    free(res), free(rest), free(eps);
}

static void vcycle(const orc_solver* s, const csc_t* A, const double* b, double* x, int K, int k) {
    cycle(s, A, b, x, K, k, s->cycle_type);
}

void orc_set_cycle_type(orc_solver* s, int cycle_type) { s->cycle_type = cycle_type; }
void orc_set_row_product(orc_solver* s, int mode) { s->row_product = mode; }

/* One V-cycle from level 0 after orc_setup (for cycle-level parity checks). */
int orc_vcycle(orc_solver* s, const int* cp, const int* ri, const double* v, const double* b, double* x, int K) {
    if (!s->have_setup) return 1;
    csc_t A = {s->n, s->n, (int*)cp, (int*)ri, (double*)v};
    vcycle(s, &A, b, x, K, 0);
    return 0;
}

/* Coarsest-level direct solve alone (after orc_setup). b, x: n_c x K column-major. */
int orc_coarse_solve(orc_solver* s, const double* b, double* x, int K) {
    if (!s->have_setup) return 1;
    const int nc = s->coarse.n;
    double* work = (double*)malloc(sizeof(double) * (size_t)(nc > 0 ? nc : 1));
    for (int c = 0; c < K; ++c) ldl_solve(&s->coarse, b + (size_t)c * nc, x + (size_t)c * nc, work);
    free(work);
    return 0;
}

/* solve(), solverType == 2 (multigrid_solver.cpp:1367-1449). x holds the initial guess on
 * entry (the binding passes rhs, core.cpp:69). hist_ms / hist_res need max_iter slots. */
int orc_solve(orc_solver* s, const int* cp, const int* ri, const double* v, const double* rhs, double* x, int K,
              double* hist_ms, double* hist_res) {
    const double t0 = now_ms();
    const int status = orc_setup(s, cp, ri, v);
    if (status) return status;
    csc_t A = {s->n, s->n, (int*)cp, (int*)ri, (double*)v};
    const double t1 = now_ms();
    int iter = 0;
    double residue;
    do {
        vcycle(s, &A, rhs, x, K, 0);
        residue = (s->smoother == 1 && s->row_product)
                      ? orc_residual_check_diff(s->n, cp, ri, v, rhs, x, K, s->criterion, s->mass)
                      : orc_residual_check(s->n, cp, ri, v, rhs, x, K, s->criterion, s->mass);
        if (hist_ms) hist_ms[iter] = now_ms() - t1;
        if (hist_res) hist_res[iter] = residue;
        ++iter;
    } while (residue > s->tol && iter < s->max_iter);
    const double t2 = now_ms();
    s->t_cycles_ms = t2 - t1;
    s->t_total_ms = t2 - t0;
    s->iterations = iter;
    s->residue = residue;
    return 0;
}

/* ------------------------------------------------------------------ accessors */
int orc_num_levels(const orc_solver* s) { return s->n_levels; }

int orc_level_shape(const orc_solver* s, int level, int* n, int* nnz) {
    if (!s->have_setup || level < 1 || level > s->n_levels) return 1;
    *n = s->Abar[level].cols;
    *nnz = s->Abar[level].colptr[s->Abar[level].cols];
    return 0;
}

int orc_get_level(const orc_solver* s, int level, int* cp, int* ri, double* v) {
    if (!s->have_setup || level < 1 || level > s->n_levels) return 1;
    const csc_t* m = &s->Abar[level];
    const int nnz = m->colptr[m->cols];
    memcpy(cp, m->colptr, sizeof(int) * (size_t)(m->cols + 1));
    memcpy(ri, m->rowidx, sizeof(int) * (size_t)nnz);
    memcpy(v, m->vals, sizeof(double) * (size_t)nnz);
    return 0;
}

/* which: 0 reduction, 1 coarsest_solve, 2 cycles, 3 solver_total, 4 iterations, 5 residue */
double orc_get_timing(const orc_solver* s, int which) {
    switch (which) {
        case 0: return s->t_reduction_ms;
        case 1: return s->t_coarse_ms;
        case 2: return s->t_cycles_ms;
        case 3: return s->t_total_ms;
        case 4: return (double)s->iterations;
        case 5: return s->residue;
    }
    return 0.0;
}
