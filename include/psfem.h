/* psfem.h -- C ABI of libpsfem.so, the B200 (sm_100a) plane-stress FEM hot path.
 *
 * The reference (Mcbr4v0/Plane-stress-FEM-with-python) is pure Python and has NO FFI; the
 * boundary it exposes is the function API of Polygones.py plus the arithmetic of its driver
 * scripts.  Each entry point below names the reference interface (file:line, relative to the
 * reference root) whose work it replaces.  The Python facade in
 * plane-stress-fem-with-python_b200/ binds exactly these symbols with ctypes.
 *
 * Conventions
 *   - every pointer named d_* is DEVICE memory owned by the caller (the Python side allocates
 *     it through torch); h_* pointers are host memory; the library keeps no caller memory;
 *   - work space is caller-allocated, sized by the matching *_workspace query;
 *   - every call takes the CUDA stream to run on (a cudaStream_t passed as void*);
 *     calls that return host scalars synchronise that stream, all others are asynchronous;
 *   - indices are int32 (a single GPU holds < 2^31 non-zeros), values are fp64;
 *   - return value: PSFEM_OK or an error code; text via psfem_last_error() (thread-local).
 */
#ifndef PSFEM_H
#define PSFEM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    PSFEM_OK = 0,
    PSFEM_EINVAL = 1,     /* bad argument */
    PSFEM_ECUDA = 2,      /* CUDA runtime error */
    PSFEM_EJACOBIAN = 3,  /* |detJ| < 1e-10: Polygones.py:686-687 ValueError; element id returned */
    PSFEM_ENOCONV = 4     /* PCG hit maxit before reaching rtol */
};

/* element kinds: (element_type, mesh_type) of Polygones.py:62 */
enum {
    PSFEM_T3 = 0, /* ('tri','coarse')  3 nodes, closed form  Polygones.py:840-883, 1218-1235 */
    PSFEM_Q4 = 1, /* ('quad','coarse') 4 nodes, 3x3 Gauss    Polygones.py:566-708            */
    PSFEM_T6 = 2, /* ('tri','fine')    6 nodes, 3-pt rule (weights sum to 1, as the reference) */
    PSFEM_Q9 = 3  /* ('quad','fine')   9 nodes, 3x3 Gauss                                      */
};

int psfem_version(void);
const char *psfem_last_error(void);
int psfem_nodes_per_element(int etype);

/* ---- mesh: Polygones.mesh, Polygones.py:62-108 ------------------------------------------
 * Nodes of the node columns [col0, col0+ncols) and elements of the cell columns
 * [cell0, cell0+ncells) of the structured mesh mesh(L,l,N,M,offX,offY,etype).  Node ids written
 * to d_conn are LOCAL: global id - col0*m (m = nodes per column), so a strip is a self-contained
 * mesh; col0=0, ncols=n, cell0=0, ncells=N gives the reference's whole mesh.
 * L_is_int / l_is_int: Python forms i*L exactly when L is an int (Polygones.py:77); pass 1 so the
 * product is formed in integer arithmetic before the division.
 * d_nodes: ncols*m*2 doubles, d_conn: ncells*M*(1 or 2)*npe int32. */
int psfem_mesh_rect(double L, double l, int64_t N, int64_t M, double offX, double offY, int etype,
                    int64_t col0, int64_t ncols, int64_t cell0, int64_t ncells,
                    double *d_nodes, int32_t *d_conn, void *stream);

/* ---- symbolic CSR pattern: replaces the dense np.zeros((2n,2n)) of Polygones.py:715,1349 ----
 * Structural pattern of K and M: row 2a+p has the sorted columns {2b+q : node b shares an
 * element with node a}.  Built by a radix sort + run-length unique of the npe^2 (a,b) node-pair
 * keys of every element.  The stable sort also yields the ASSEMBLY PLAN: for every unique pair
 * the contributing (element,i,j) codes in ascending element id -- the order of the reference's
 * scatter loop Polygones.py:717-726 -- which makes the numeric phase a deterministic, atomic-free
 * segmented reduction.
 *   step 1  psfem_pattern_count : sort/unique; returns n_pairs (nnz = 4*n_pairs); state in ws
 *   step 2  psfem_pattern_fill  : writes row_ptr, col_idx and the plan into caller arrays
 */
int psfem_pattern_workspace(int64_t nel, int npe, int64_t n_nodes, size_t *bytes);
int psfem_pattern_count(const int32_t *d_conn, int64_t nel, int npe, int64_t n_nodes,
                        void *d_ws, size_t ws_bytes, int64_t *h_n_pairs, void *stream);
int psfem_pattern_fill(int64_t nel, int npe, int64_t n_nodes, int64_t n_pairs,
                       const void *d_ws, size_t ws_bytes,
                       int32_t *d_row_ptr,   /* 2*n_nodes+1 */
                       int32_t *d_col_idx,   /* 4*n_pairs   */
                       int32_t *d_node_ptr,  /* n_nodes+1 : first pair of each node row        */
                       int32_t *d_pair_row,  /* n_pairs   : row node of each pair               */
                       int32_t *d_pair_ptr,  /* n_pairs+1 : segment bounds into d_perm          */
                       int32_t *d_perm,      /* nel*npe^2 : contribution codes e*npe^2+i*npe+j  */
                       void *stream);

/* ---- numeric assembly: Polygones.global_stiffness_matrix :710-728 / global_mass_matrix :1344-1364
 * what: bit0 = K (element_stiffness_matrix :566-655, T3 :875-883), bit1 = M (element_mass_matrix
 * :1250-1342, T3 :1228-1235).  d_K_vals / d_M_vals: nnz doubles on the pattern above (M keeps
 * explicit zeros at u-v couplings).  h_bad_element: first element (lowest id) with |detJ|<1e-10,
 * else -1; the call then returns PSFEM_EJACOBIAN.  Synchronises the stream. */
int psfem_assemble_workspace(int64_t nel, int etype, int what, size_t *bytes);
int psfem_assemble(const double *d_nodes, const int32_t *d_conn, int64_t nel, int etype, int64_t n_nodes,
                   double E, double nu, double rho, double t, int what,
                   int64_t n_pairs, const int32_t *d_node_ptr, const int32_t *d_pair_row,
                   const int32_t *d_pair_ptr, const int32_t *d_perm,
                   double *d_K_vals, double *d_M_vals,
                   void *d_ws, size_t ws_bytes, int64_t *h_bad_element, void *stream);

/* element matrices only (block layout e-major): Polygones.element_stiffness_matrix :566,
 * element_mass_matrix :1250.  d_Ke: nel*(2npe)^2 row-major per element; d_Me likewise. */
int psfem_element_matrices(const double *d_nodes, const int32_t *d_conn, int64_t nel, int etype,
                           double E, double nu, double rho, double t, int what,
                           double *d_Ke, double *d_Me, int64_t *h_bad_element, void *stream);

/* ---- boundary conditions: Polygones.apply_boundary_conditions :989-1006, reduce :898-911 ----
 * psfem_free_map: d_constrained (nc DOF ids, duplicates allowed) -> d_new_id[n] (reduced index or
 * -1) and d_free_dofs (ascending complement, Polygones.py:1000); returns n_free.
 * psfem_reduce_count/fill: K[np.ix_(free,free)] on CSR; d_keep[k] = position in the full value
 * array of reduced entry k, so M and K_eff reduce with psfem_gather. */
int psfem_free_map_workspace(int64_t n, size_t *bytes);
int psfem_free_map(int64_t n, const int32_t *d_constrained, int64_t nc, int32_t *d_new_id,
                   int32_t *d_free_dofs, int64_t *h_n_free, void *d_ws, size_t ws_bytes, void *stream);
int psfem_reduce_count(int64_t n, const int32_t *d_row_ptr, const int32_t *d_col_idx, const int32_t *d_new_id,
                       int64_t n_free, int32_t *d_row_ptr_red, int64_t *h_nnz_red,
                       void *d_ws, size_t ws_bytes, void *stream);
int psfem_reduce_fill(int64_t n, const int32_t *d_row_ptr, const int32_t *d_col_idx, const int32_t *d_new_id,
                      const int32_t *d_row_ptr_red, int32_t *d_col_idx_red, int32_t *d_keep, void *stream);
int psfem_gather(const double *d_src, const int32_t *d_index, int64_t n, double *d_dst, void *stream);
/* reconstruct_full_vector(s) Polygones.py:913-936: d_full[free[i]] = d_red[i], zeros elsewhere (nvec rows) */
int psfem_scatter_free(const double *d_red, const int32_t *d_free_dofs, int64_t n_free, int64_t n_full,
                       int64_t nvec, double *d_full, void *stream);

/* ---- sparse kernels used by the static / Newmark / modal drivers -------------------------
 * SpMV stages row blocks of the CSR stream through shared memory (coalesced vals/col_idx reads,
 * per-row reduction from shared memory); d_rowblk from psfem_spmv_plan. */
int psfem_spmv_plan_size(int64_t n, int64_t nnz, int64_t *n_blocks);
int psfem_spmv_plan(int64_t n, const int32_t *d_row_ptr, int64_t nnz, int32_t *d_rowblk, int64_t n_blocks, void *stream);
/* y = alpha*A*x + beta*z   (z may be NULL -> beta ignored; y may alias z) */
int psfem_spmv(int64_t n, const int32_t *d_row_ptr, const int32_t *d_col_idx, const double *d_vals,
               const int32_t *d_rowblk, int64_t n_blocks,
               const double *d_x, double alpha, double beta, const double *d_z, double *d_y, void *stream);
/* y = alpha*A1*x1 + beta*A2*x2 + gamma*z on a shared pattern (Newmark RHS main.py:310) */
int psfem_spmv2(int64_t n, const int32_t *d_row_ptr, const int32_t *d_col_idx,
                const double *d_vals1, const double *d_vals2, const int32_t *d_rowblk, int64_t n_blocks,
                const double *d_x1, const double *d_x2, double alpha, double beta, double gamma,
                const double *d_z, double *d_y, void *stream);
int psfem_extract_diag(int64_t n, const int32_t *d_row_ptr, const int32_t *d_col_idx, const double *d_vals,
                       int64_t col_offset, double *d_diag, void *stream);

/* Jacobi-preconditioned CG: replaces np.linalg.solve / scipy.linalg.solve / lu_solve of
 * Statictest.py:69, StaticTest2.py:71, Beamexemple.py:63, main.py:286,314.
 * x0 = d_x on entry, stop when ||r||_2 <= rtol*||b||_2, residual checked every check_every
 * iterations (device-side scalars; one host sync per check).  h_info: [0]=iterations,
 * [1]=relative residual, [2]=b norm. */
int psfem_pcg_workspace(int64_t n, int64_t n_blocks, size_t *bytes);
int psfem_pcg_jacobi(int64_t n, const int32_t *d_row_ptr, const int32_t *d_col_idx, const double *d_vals,
                     const int32_t *d_rowblk, int64_t n_blocks,
                     const double *d_b, double *d_x, double rtol, int64_t maxit, int64_t check_every,
                     double *h_info, void *d_ws, size_t ws_bytes, void *stream);

/* The three kernels of one PCG iteration, exposed so that the strip-partitioned multi-GPU
 * driver can put its halo exchange and all-reduces between them.  Scalars live in d_scal:
 * [0]=rz [1]=pq [2]=rz_new [3]=rr [4]=alpha [5]=beta.  col indices address the EXTENDED local
 * vector (owned rows start at x_offset). */
int psfem_pcg_spmv_dot(int64_t n_rows, const int32_t *d_row_ptr, const int32_t *d_col_idx, const double *d_vals,
                       const int32_t *d_rowblk, int64_t n_blocks, const double *d_p_ext, int64_t x_offset,
                       double *d_q, double *d_scal, double *d_partials, void *stream);
int psfem_pcg_update(int64_t n, double *d_x, double *d_r, const double *d_p, const double *d_q,
                     const double *d_dinv, double *d_scal, double *d_partials, void *stream);
int psfem_pcg_pupdate(int64_t n, double *d_p, const double *d_r, const double *d_dinv,
                      const double *d_scal, void *stream);
int psfem_pcg_init(int64_t n, const double *d_r, const double *d_dinv, double *d_p, double *d_scal,
                   double *d_partials, void *stream);

/* ---- BLAS-1 helpers for the drivers (Lanczos, energies) ---------------------------------- */
int psfem_axpby(int64_t n, double a, const double *d_x, double b, const double *d_y, double *d_out, void *stream);
int psfem_dot(int64_t n, const double *d_x, const double *d_y, double *h_out, void *d_ws, size_t ws_bytes, void *stream);
int psfem_dot_workspace(int64_t n, size_t *bytes);
/* out[j] = V[j,:] . w  for j<m  (V row-major m x n): Gram-Schmidt coefficients */
int psfem_multi_dot(int64_t m, int64_t n, const double *d_V, const double *d_w, double *d_out, void *stream);
/* w -= sum_j c[j] V[j,:] */
int psfem_multi_axpy(int64_t m, int64_t n, const double *d_V, const double *d_c, double *d_w, void *stream);

/* ---- Newmark-beta as implemented by main.py:264-332 (NOT the textbook scheme) -------------
 * Runs steps 1..n_steps-1 on the reduced system.  Load of step i: d_fshape * h_fscale[i]
 * (main.py:226-251 is a fixed nodal pattern times a time ramp).  K, M, K_eff share the pattern.
 * d_U, d_V, d_A: n_store columns each (column-major, n rows), column c holds step c*store_every.
 * d_energy: n_steps x 3 (kinetic, potential, total; main.py:377-385) or NULL.
 * Solves use Jacobi-PCG with rtol, warm-started from the previous acceleration.
 * h_info: [0]=total PCG iterations, [1]=max relres, [2]=number of stabiliser triggers (main.py:324-332). */
int psfem_newmark_workspace(int64_t n, int64_t n_blocks, size_t *bytes);
int psfem_newmark_run(int64_t n, const int32_t *d_row_ptr, const int32_t *d_col_idx,
                      const double *d_K, const double *d_M, const double *d_Keff,
                      const int32_t *d_rowblk, int64_t n_blocks,
                      const double *d_fshape, const double *h_fscale, int64_t n_steps,
                      double dt, double gamma, double beta, double alpha_R, double beta_R,
                      const double *d_a0, int64_t store_every,
                      double *d_U, double *d_V, double *d_A, double *d_energy,
                      double rtol, int64_t maxit, double *h_info,
                      void *d_ws, size_t ws_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PSFEM_H */
