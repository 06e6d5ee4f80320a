/*
 * hx_b200.h -- C-ABI of libhx_b200.so, the B200 (sm_100a) replacement for the
 * third-party routines helmholtz-x reaches on its hot path.
 *
 * The reference (ekremekc/helmholtz-x) is pure Python and has no FFI of its own:
 * its boundary is the helmholtz_x Python API whose objects are DOLFINx / PETSc /
 * SLEPc handles.  Each entry point below replaces the DOLFINx/PETSc/SLEPc routine
 * that the cited reference line calls; the Python shim in helmholtz_x_b200/ keeps
 * the reference's class and function names and binds these symbols with ctypes
 * (INTEGRATION.md shows the stub).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in _h (host);
 *  - complex128 is interleaved (re,im) doubles (torch.complex128 / PETSc complex);
 *  - CSR: int32 row pointers (n+1), int32 column indices, sorted, no duplicates;
 *  - the caller owns every buffer (torch caching allocator); the library keeps no
 *    state besides read-only constant tables; `stream` is a cudaStream_t;
 *  - return 0 on success, negative on error; hx_last_error() gives the text;
 *  - nothing here falls back to the CPU: without a CUDA device every compute
 *    entry point returns HX_ERR_CUDA.
 */
#ifndef HX_B200_H
#define HX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HX_OK 0
#define HX_ERR_CUDA (-1)
#define HX_ERR_ARG (-2)
#define HX_ERR_CAPACITY (-3)

typedef void* hx_stream_t; /* cudaStream_t */

const char* hx_last_error(void);
int hx_version(void);
/* number of kernel launches issued through this library since the last reset
 * (bench.py reports it as gpu_launches) */
int64_t hx_launch_count(void);
void hx_launch_count_reset(void);
/* kernels run by a CUDA-graph replay of captured hx_* launches do not pass through the entry
 * points; the host adds them here so the counter stays the number of kernels executed */
void hx_launch_count_add(int64_t n);

/* ------------------------------------------------------------------ K7 / K8
 * PETSc MatMult on AIJ (helmholtz_x/petsc4py_utils.py:86,96; every SLEPc
 * operator apply).  y = alpha * M x + beta * y0   (y0 may be NULL => beta term
 * dropped; y0 may alias y).  lanes = threads cooperating on one row
 * (2,4,8,16,32) or 0 = choose from nnz/n. */
int hx_spmv_zz(int n, const int32_t* indptr, const int32_t* indices, const double* vals_c128,
               const double* x_c128, double* y_c128, const double* alpha_c128_h,
               const double* beta_c128_h, const double* y0_c128, int lanes, hx_stream_t stream);
/* same with real-valued matrix entries (A, C are real: 12 B/nnz instead of 20) */
int hx_spmv_dz(int n, const int32_t* indptr, const int32_t* indices, const double* vals_f64,
               const double* x_c128, double* y_c128, const double* alpha_c128_h,
               const double* beta_c128_h, const double* y0_c128, int lanes, hx_stream_t stream);
/* SELL-32-sigma variant (rows sorted by length in windows, 32-row slices stored
 * column-major: fully coalesced, one thread per row).  y[row_perm[r]] = ...;
 * alpha/beta/y0 as in hx_spmv_zz (all NULL => y = M x).  variant: 0 = default tuning,
 * 1..5 = (unroll, min CTAs/SM) alternatives kept for the roofline study. */
int hx_spmv_sell_zz(int n, int n_slices, const int64_t* slice_ptr, const int32_t* cols,
                    const double* vals_c128, const int32_t* row_perm, const double* x_c128,
                    double* y_c128, const double* alpha_c128_h, const double* beta_c128_h,
                    const double* y0_c128, int variant, hx_stream_t stream);
/* damped-Jacobi sweep on the SELL matrix: xout = xin + omega * dinv .* (b - M xin) */
int hx_jacobi_sell(int n, int n_slices, const int64_t* slice_ptr, const int32_t* cols,
                   const double* vals_c128, const int32_t* row_perm, const double* dinv_c128,
                   const double* b_c128, const double* xin_c128, double* xout_c128, double omega,
                   int variant, hx_stream_t stream);
/* CSR -> SELL-32 conversion: slice widths, then fill (sell_vals and/or the
 * sell-position -> csr-position map `src` may be NULL), then value refresh by gather */
int hx_sell_slice_widths(int n, const int32_t* indptr, const int32_t* row_perm, int n_slices,
                         int32_t* widths, hx_stream_t stream);
int hx_sell_fill(int n, const int32_t* indptr, const int32_t* indices, const double* vals_c128,
                 const int32_t* row_perm, int n_slices, const int64_t* slice_ptr, int32_t* cols,
                 double* sell_vals_c128, int32_t* src, hx_stream_t stream);
int hx_sell_gather(int64_t total, const int32_t* src, const double* csr_vals_c128, double* sell_vals_c128,
                   hx_stream_t stream);

/* complex64 variants of the same kernels: the multigrid V-cycle used as GMRES
 * preconditioner runs in single precision (matrix 12 B/nnz instead of 20), the outer
 * GMRES / Krylov-Schur stay complex128.  Pointers are float (interleaved re,im). */
int hx_spmv_cc(int n, const int32_t* indptr, const int32_t* indices, const float* vals_c64,
               const float* x_c64, float* y_c64, const double* alpha_c128_h, const double* beta_c128_h,
               const float* y0_c64, int lanes, hx_stream_t stream);
int hx_spmv_sc(int n, const int32_t* indptr, const int32_t* indices, const float* vals_f32,
               const float* x_c64, float* y_c64, const double* alpha_c128_h, const double* beta_c128_h,
               const float* y0_c64, int lanes, hx_stream_t stream);
int hx_spmv_sell_cc(int n, int n_slices, const int64_t* slice_ptr, const int32_t* cols,
                    const float* vals_c64, const int32_t* row_perm, const float* x_c64, float* y_c64,
                    const double* alpha_c128_h, const double* beta_c128_h, const float* y0_c64,
                    int variant, hx_stream_t stream);
int hx_jacobi_sell_c(int n, int n_slices, const int64_t* slice_ptr, const int32_t* cols,
                     const float* vals_c64, const int32_t* row_perm, const float* dinv_c64,
                     const float* b_c64, const float* xin_c64, float* xout_c64, double omega,
                     int variant, hx_stream_t stream);
int hx_jacobi_sweep_c(int n, const int32_t* indptr, const int32_t* indices, const float* vals_c64,
                      const float* dinv_c64, const float* b_c64, const float* xin_c64, float* xout_c64,
                      double omega, int lanes, hx_stream_t stream);
/* real float32 matrix (the prolongation / restriction operators of the cycle) times complex64 vector in
 * SELL-32, and the matching value gather */
int hx_spmv_sell_sc(int n, int n_slices, const int64_t* slice_ptr, const int32_t* cols,
                    const float* vals_f32, const int32_t* row_perm, const float* x_c64, float* y_c64,
                    const double* alpha_c128_h, const double* beta_c128_h, const float* y0_c64,
                    int variant, hx_stream_t stream);
int hx_sell_gather_s(int64_t total, const int32_t* src, const float* csr_vals_f32, float* sell_vals_f32,
                     hx_stream_t stream);
/* SELL value refresh with down-conversion complex128 CSR values -> complex64 SELL values */
int hx_sell_gather_c(int64_t total, const int32_t* src, const double* csr_vals_c128, float* sell_vals_c64,
                     hx_stream_t stream);

/* K8: P(sigma) = A + sigma B + sigma^2 C on the shared cell pattern, replacing the
 * MatAXPY chain of helmholtz_x/eigensolvers.py:174-176,240,309-315.  A, C real on
 * the full pattern; B complex on the same pattern or NULL.  out complex.
 * coefficients: out = ca*A + cb*B + cc*C (all complex, host). */
int hx_combine_abc(int64_t nnz, const double* a_f64, const double* b_c128, const double* c_f64,
                   const double* ca_h, const double* cb_h, const double* cc_h, double* out_c128,
                   hx_stream_t stream);
/* Same with A, B, C all complex on one pattern (any may be NULL): the Bloch-reduced operators
 * NB*A*BN, NB*C*BN of helmholtz_x/bloch_operator.py:70-78 are Hermitian, not real. */
int hx_combine_zzz(int64_t nnz, const double* a_c128, const double* b_c128, const double* c_c128,
                   const double* ca_h, const double* cb_h, const double* cc_h, double* out_c128,
                   hx_stream_t stream);
/* matrix-free flame term (flame_matrices.py:75-108 without the outer product):
 * t_f = sum_k rval[k] x[ridx[k]] over k in [rptr[f], rptr[f+1]);
 * y[lrow[i]] += coef * sum_{k in [lptr[i],lptr[i+1])} lval[k] * t[lcol[k]]  */
int hx_lowrank_dots(int r, const int32_t* rptr, const int32_t* ridx, const double* rval_f64,
                    const double* x_c128, double* t_c128, hx_stream_t stream);
int hx_lowrank_update(int nrows, const int32_t* lrow, const int32_t* lptr, const int32_t* lcol,
                      const double* lval_f64, const double* t_c128, const double* coef_c128_h,
                      double* y_c128, hx_stream_t stream);

/* ------------------------------------------------------------------ K10 / K11
 * SLEPc BVOrthogonalize / VecDot / VecNorm / BVMultInPlace on an n x k basis
 * stored as k contiguous vectors with leading dimension ld (complex elements).
 * Reductions are two-stage with a fixed tree: bitwise run-to-run reproducible.
 * scratch: at least hx_reduce_scratch_bytes(k) bytes. */
int64_t hx_reduce_scratch_bytes(int k);
/* out[j] = sum_i conj(V[j][i]) * w[i]  (conj=1)  or  sum_i V[j][i]*w[i] (conj=0) */
int hx_multi_dot(int64_t n, int k, const double* V_c128, int64_t ld, const double* w_c128,
                 int conj, double* out_c128, void* scratch, hx_stream_t stream);
/* w -= sum_j h[j] V[j];  hacc[j] += h[j] (hacc may be NULL);
 * nrm2_out (device, 1 double, may be NULL) = ||w_new||^2 */
int hx_multi_axpy(int64_t n, int k, const double* V_c128, int64_t ld, const double* h_c128,
                  double* w_c128, double* hacc_c128, double* nrm2_out, void* scratch,
                  hx_stream_t stream);
/* out = w * (1/sqrt(*nrm2_dev))  (or  w * alpha_h if nrm2_dev NULL) */
int hx_scale_copy(int64_t n, const double* w_c128, const double* nrm2_dev, const double* alpha_c128_h,
                  double* out_c128, hx_stream_t stream);
/* y = a*x + b*y  (a,b complex host scalars) */
int hx_axpby(int64_t n, const double* a_h, const double* x_c128, const double* b_h, double* y_c128,
             hx_stream_t stream);
/* Vout[c] = sum_j Q[j + c*ldq] V[j], c<kout  (restart V <- V Q; Q column-major m x kout, device) */
int hx_basis_rotate(int64_t n, int m, int kout, const double* V_c128, int64_t ld, const double* Q_c128,
                    int ldq, double* Vout_c128, int64_t ldout, hx_stream_t stream);

/* the same contraction on the FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64; four real products per
 * complex k-step, 16 output vectors per pass over V) -- SLEPc BVMultInPlace at the Krylov-Schur restart */
int hx_basis_rotate_dmma(int64_t n, int m, int kout, const double* V_c128, int64_t ld, const double* Q_c128,
                         int ldq, double* Vout_c128, int64_t ldout, hx_stream_t stream);

/* ------------------------------------------------------------------ K9
 * inner solve building blocks (what PETSc KSP+PC LU/MUMPS did behind ST sinvert) */
/* xout = xin + omega * dinv .* (b - M xin); xin NULL => xout = omega*dinv.*b */
int hx_jacobi_sweep(int n, const int32_t* indptr, const int32_t* indices, const double* vals_c128,
                    const double* dinv_c128, const double* b_c128, const double* xin_c128,
                    double* xout_c128, double omega, int lanes, hx_stream_t stream);
int hx_extract_diag_inv(int n, const int32_t* indptr, const int32_t* indices, const double* vals_c128,
                        double* dinv_c128, hx_stream_t stream);
/* ILU(0) on the CSR pattern, level-scheduled: rows of level l are
 * level_rows[level_ptr[l]..level_ptr[l+1]) (host array level_ptr_h). */
int hx_ilu0_factor(int n, const int32_t* indptr, const int32_t* indices, const int32_t* diag_pos,
                   double* lu_c128, int n_levels, const int32_t* level_ptr_h, const int32_t* level_rows,
                   hx_stream_t stream);
int hx_ilu0_solve(int n, const int32_t* indptr, const int32_t* indices, const int32_t* diag_pos,
                  const double* lu_c128, int n_levels, const int32_t* level_ptr_h, const int32_t* level_rows,
                  int n_levels_u, const int32_t* level_ptr_u_h, const int32_t* level_rows_u,
                  const double* b_c128, double* x_c128, hx_stream_t stream);
int hx_diag_positions(int n, const int32_t* indptr, const int32_t* indices, int32_t* diag_pos, hx_stream_t stream);
/* dense coarse-level solve: in-place Gauss-Jordan inverse with partial pivoting
 * (single CTA, n <= 1024), then y = Ainv x */
int hx_dense_inverse(int n, double* a_c128_colmajor, int32_t* info_dev, hx_stream_t stream);
int hx_dense_gemv(int n, const double* a_c128_colmajor, const double* x_c128, double* y_c128, hx_stream_t stream);

/* The tail of the multigrid cycle as ONE persistent kernel (grid-wide barriers instead of kernel boundaries):
 * on the last smoothed level (complex64 CSR, n rows) nu pre-sweeps (the first from zero), residual, restriction
 * (float32 CSR, nc rows), dense coarsest solve xc = Ainv bc (complex128 column-major inverse), prolongation
 * (float32 CSR, n rows), nu post-sweeps.  omega[s] = damping of sweep s (pre and post alike).  The result ends in
 * buf0 when the number of buffer swaps 2 nu - 1 is even, else in buf1 (the host knows nu).
 * barrier: device uint32[2], zero-initialised, private to this descriptor. */
typedef struct {
    int32_t n, nc, nu, pad_;
    const int32_t* a_ptr; const int32_t* a_idx; const float* a_val;      /* level operator, complex64 values */
    const float* dinv;                                                   /* complex64 */
    const int32_t* r_ptr; const int32_t* r_idx; const float* r_val;      /* restriction, float32 */
    const int32_t* p_ptr; const int32_t* p_idx; const float* p_val;      /* prolongation, float32 */
    const double* coarse_inv;                                            /* complex128, nc x nc column-major */
    float omega[4];
    void* buf0; void* buf1; void* r;                                     /* complex64, n each */
    void* bc; void* xc;                                                  /* complex128, nc each */
    void* barrier;
} hx_tail_desc;
int hx_amg_tail(const hx_tail_desc* desc_h, const float* b_c64, hx_stream_t stream);

/* multigrid set-up: C = A * B on CSR (real double), one warp per row.  symbolic: distinct
 * column count per row (write_cols=0; -1 = more than 2048 columns, caller falls back) or the
 * sorted columns (write_cols=1); numeric: deterministic accumulation into a given pattern. */
int hx_spgemm_symbolic(int m, const int32_t* a_ptr, const int32_t* a_idx, const int32_t* b_ptr,
                       const int32_t* b_idx, int32_t* row_nnz, const int32_t* c_ptr, int32_t* c_idx,
                       int write_cols, hx_stream_t stream);
int hx_spgemm_numeric(int m, const int32_t* a_ptr, const int32_t* a_idx, const double* a_val,
                      const int32_t* b_ptr, const int32_t* b_idx, const double* b_val,
                      const int32_t* c_ptr, const int32_t* c_idx, double* c_val, hx_stream_t stream);
/* size classes: hx_spgemm_symbolic with write_cols 2 / 3 (count / fill) and hx_spgemm_numeric_small serve rows of at
 * most 512 distinct columns with 8 warps per CTA (the fine levels); overflow is reported as -1 in row_nnz and the
 * caller repeats with the large class */
int hx_spgemm_numeric_small(int m, const int32_t* a_ptr, const int32_t* a_idx, const double* a_val,
                            const int32_t* b_ptr, const int32_t* b_idx, const double* b_val,
                            const int32_t* c_ptr, const int32_t* c_idx, double* c_val, hx_stream_t stream);

/* ------------------------------------------------------------------ K4
 * DOLFINx SparsityPattern + MatCreateAIJ (acoustic_matrices.py:102): CSR pattern
 * from the cell dofmap (n_cells x nd, int32).  Three steps so the caller can
 * allocate between them. */
int hx_dof_cell_count(int64_t n_cells, int nd, const int32_t* cell_dofs, int n_dofs, int32_t* count,
                      hx_stream_t stream);
int hx_dof_cell_fill(int64_t n_cells, int nd, const int32_t* cell_dofs, int n_dofs, const int32_t* adj_ptr,
                     int32_t* cursor, int32_t* adj_cells, hx_stream_t stream);
/* per-row unique column count (write_cols=0) or fill (write_cols=1) */
int hx_pattern_rows(int n_dofs, int nd, const int32_t* cell_dofs, const int32_t* adj_ptr, int32_t* adj_cells,
                    int32_t* row_nnz, const int32_t* indptr, int32_t* indices, int write_cols,
                    hx_stream_t stream);
/* Colouring of cells (or facets) so that entities of one colour share no vertex, on the device:
 * Jones-Plassmann rounds with fixed hashed priorities; decisions read only the state at the start of
 * a round, so the result depends on the mesh alone (bitwise reproducible assembly).
 * adj_ptr/adj_cells: vertex -> entity adjacency (hx_dof_cell_count / hx_dof_cell_fill with nd = nv);
 * color_a (initialised to -1) and color_b ping-pong, `rounds` (even) rounds per call, result in
 * color_a; *remaining_dev = entities still uncoloured (call again until 0).  -2 = more than 128 colours. */
int hx_color_cells(int64_t n_cells, int nv, const int32_t* cells, const int32_t* adj_ptr,
                   const int32_t* adj_cells, int32_t* color_a, int32_t* color_b, int rounds,
                   uint64_t* remaining_dev, hx_stream_t stream);
/* the same job as a host greedy first-fit (kept as the checker of the device colouring in tests;
 * arrays are HOST pointers). returns #colours or <0 */
int hx_color_cells_h(int64_t n_cells, int nv, const int32_t* cells_h, int n_nodes, int32_t* color_h);

/* ------------------------------------------------------------------ K1 / K2 / K3
 * FFCx tabulate_tensor + fem::assemble_matrix (acoustic_matrices.py:101-103,
 * 121-123): per-cell P1/P2 tetrahedral element matrices scattered colour by
 * colour (cells[color_ptr_h[c]..color_ptr_h[c+1]) of color_cells) into real CSR
 * values a_vals (A = -int c^2 grad.grad) and c_vals (C = int phi phi).
 * c_field: P1 nodal (c_is_dg0=0) or per cell. */
int hx_assemble_AC(int degree, int64_t n_cells, const double* x_f64, const int32_t* cells,
                   const int32_t* cell_dofs, const double* c_field, int c_is_dg0, int n_colors,
                   const int64_t* color_ptr_h, const int32_t* color_cells, const int32_t* indptr,
                   const int32_t* indices, double* a_vals, double* c_vals, hx_stream_t stream);
/* exterior-facet term (acoustic_matrices.py:71,83,95,108-109):
 * B += coef * int c phi_k phi_j ds over the listed facets (coef = i/Z, host). */
int hx_assemble_B(int degree, int64_t n_facets, const double* x_f64, const int32_t* facets,
                  const int32_t* facet_dofs, const int32_t* facet_cell, const double* c_field, int c_is_dg0,
                  const double* coef_c128_h, int n_colors, const int64_t* color_ptr_h,
                  const int32_t* color_facets, const int32_t* indptr, const int32_t* indices,
                  double* b_vals_c128, hx_stream_t stream);
/* Dirichlet rows/cols -> 0, diagonal -> 1 (DOLFINx assemble_matrix(bcs=...)) */
int hx_apply_dirichlet(int n, const int32_t* indptr, const int32_t* indices, const uint8_t* is_bc,
                       double* vals, int is_complex, hx_stream_t stream);
/* sum over listed facets of area and area*mean(f) (choked BCs, acoustic_matrices.py:76-90)
 * out[0]=area, out[1]=int f ds */
int hx_facet_integrals(int64_t n_facets, const double* x_f64, const int32_t* facets, const double* f_nodal,
                       double* out2, void* scratch, hx_stream_t stream);
/* cell volumes (Q_multiple / normalize, parameters_utils.py:228-247) */
int hx_cell_volumes(int64_t n_cells, const double* x_f64, const int32_t* cells, double* vol, hx_stream_t stream);

/* ------------------------------------------------------------------ K5 / K6
 * assemble_vector for the flame forms (flame_matrices.py:141,199-200).
 * left: out[dof] += scale * int g h phi, g = (gamma-1) nodal (gm1_nodal) or the
 *       constant gm1_const when gm1_nodal NULL; h P1 nodal (h_is_dg0=0) or DG0;
 *       only cells with cell_tags[c]==tag are visited when cell_tags != NULL.
 * right: out[dof] += int d(phi)/dz * w/rho  (w, rho P1 nodal; degree-2 4-point
 *       rule for P1, 14-point degree-5 rule for P2) */
int hx_flame_left(int degree, int64_t n_cells, const double* x_f64, const int32_t* cells,
                  const int32_t* cell_dofs, const double* gm1_nodal, double gm1_const, const double* h,
                  int h_is_dg0, double scale, const int32_t* cell_tags, int tag, int n_colors,
                  const int64_t* color_ptr_h, const int32_t* color_cells, double* out_f64, hx_stream_t stream);
int hx_flame_right(int degree, int64_t n_cells, const double* x_f64, const int32_t* cells,
                   const int32_t* cell_dofs, const double* w_nodal, const double* rho_nodal, int n_colors,
                   const int64_t* color_ptr_h, const int32_t* color_cells, double* out_f64, hx_stream_t stream);
/* determine_point_ownership + Expression.eval (flame_matrices.py:144-156):
 * owner[p] = lowest cell index whose barycentric coordinates of point p are all
 * >= -tol (INT_MAX if none); then d(phi_a)/dz at the point for the owner's dofs. */
int hx_locate_points(int64_t n_cells, const double* x_f64, const int32_t* cells, int n_points,
                     const double* points_f64, double tol, int32_t* owner, hx_stream_t stream);
int hx_point_dphidz(int degree, const double* x_f64, const int32_t* cells, int n_points,
                    const double* points_f64, const int32_t* owner, double* out_np_by_nd, hx_stream_t stream);
/* adjoint sensitivity tail (helmholtz_x/shape_derivatives.py:12-37):
 * out[0] = int over the listed boundary facets of (V.n) div(conj(p_adj) c^2 grad p) ds,
 * V P1 nodal (n_nodes x 3), c P1 nodal, p / p_adj complex128 in the degree-`degree` space,
 * facet_cell = owning cell of every boundary facet. */
int hx_shape_derivative(int degree, int n_sel, const int32_t* sel, const double* x_f64, const int32_t* cells,
                        const int32_t* cell_dofs, const int32_t* facets, const int32_t* facet_cell,
                        const double* V_f64, const double* p_c128, const double* p_adj_c128,
                        const double* c_nodal_f64, double* out_c128, hx_stream_t stream);
/* |v|<tol -> 0 (flame_matrices.py:67-68; values are real) */
int hx_threshold(int64_t n, double* v_f64, double tol, hx_stream_t stream);

/* ------------------------------------------------------------------ multi-GPU (row e)
 * Peer-memory data path for one process per GPU on one NVLink/NVSwitch node.  Replaces, on the
 * per-iteration path, PETSc's VecScatter inside MatMult on an MPIAIJ matrix
 * (helmholtz_x/petsc4py_utils.py:86,96 under mpirun) and the MPI_Allreduce inside
 * VecDot / VecNorm / BVOrthogonalize (helmholtz_x/eigensolvers.py:62,113): the kernels below store
 * straight into the neighbours' HBM and synchronise with sequence flags -- no NCCL call and no host
 * round trip, so they can be captured in a CUDA graph with the kernels around them.
 * (torch.distributed/NCCL remains the set-up plumbing.) */
#define HX_PEER_MAX 15          /* neighbours of one rank (world <= 16) */
#define HX_PEER_FLAG_KINDS 3    /* flag block of a rank: [HX_PEER_FLAG_KINDS][world] uint64 */

/* hx_peer_alloc: cudaMalloc + zero fill + IPC handle (64 bytes) of a buffer other ranks may map;
 * hx_peer_open maps a handle received from another process (peer access enabled lazily). */
int hx_peer_alloc(int64_t bytes, void** ptr_out_h, unsigned char* handle64_h);
int hx_peer_open(const unsigned char* handle64_h, void** ptr_out_h);
int hx_peer_close(void* ptr);
int hx_peer_free(void* ptr);

/* one neighbour exchange of interface values (host descriptor, passed to the kernel by value) */
typedef struct {
    int32_t world, rank, n_nb, pad_;
    int32_t nb_rank[HX_PEER_MAX + 1];  /* neighbour ranks */
    int64_t send_ptr[HX_PEER_MAX + 1]; /* send_idx[send_ptr[i]..send_ptr[i+1]) goes to neighbour i */
    const int32_t* send_idx;           /* device: local indices of the values to send */
    void* dst[HX_PEER_MAX];            /* neighbour i's memory: start of the ghost segment this rank fills */
    void* nb_flags[HX_PEER_MAX];       /* neighbour i's flag block (in its memory) */
    void* my_flags;                    /* this rank's flag block */
    void* chan_seq;                    /* device uint64[world]: exchanges completed per channel */
    void* block_counter;               /* device uint32 (zero) */
    int32_t* err;                      /* device int32: set non-zero when a wait timed out */
} hx_peer_halo_desc;
/* x_local[send_idx[..]] -> neighbours' ghost segments; returns (in stream order) when this rank's
 * ghost values have arrived as well.  elem_bytes: 16 (complex128) or 8 (complex64). */
int hx_peer_halo_exchange(const hx_peer_halo_desc* plan_h, const void* x_local, int elem_bytes, hx_stream_t stream);

typedef struct {
    int32_t world, rank;
    int64_t slot_bytes;                /* capacity of one contribution */
    void* slots[HX_PEER_MAX + 1];      /* rank q's slot area: [2][world][slot_bytes] */
    void* flags[HX_PEER_MAX + 1];      /* rank q's flag block */
    void* my_flags;
    void* seq;                         /* device uint64: all-reduces completed */
    void* block_counter;               /* device uint32[2] (zero) */
    int32_t* err;
} hx_peer_allreduce_desc;
/* out[i] = sum over ranks of in[i], summed in rank order (bitwise identical on every rank).
 * count doubles (is_f32=0) or floats (is_f32=1); in may alias out. */
int hx_peer_allreduce(const hx_peer_allreduce_desc* ar_h, const void* in, void* out, int64_t count, int is_f32,
                      hx_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HX_B200_H */
