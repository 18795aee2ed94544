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
/* bind the calling host thread to a CUDA device (the library links its own static CUDA runtime;
 * the Python side calls this with torch's current device before any other entry point) */
int psfem_set_device(int device);

/* CSR operand of the sparse kernels.  row_ptr values are absolute offsets into col_idx / vals, so a
 * contiguous row slice of a larger matrix (the owned rows of a strip) is a valid operand:
 * row_ptr + first_row, n_rows = rows in the slice.  col_idx address the (extended) input vector.
 * rowblk / n_blocks / max_row_nnz come from psfem_spmv_plan. */
typedef struct psfem_csr {
    int64_t n_rows;
    int64_t nnz;            /* row_ptr[n_rows] - row_ptr[0] */
    const int32_t *row_ptr; /* n_rows + 1 */
    const int32_t *col_idx;
    const double *vals;
    const int32_t *rowblk;  /* n_blocks + 2 ints (plan + one scratch slot) */
    int64_t n_blocks;
    int64_t max_row_nnz;
    /* optional node-block view (psfem_spmv_nb_plan; all NULL / 0 when the matrix has no 2x2 node-block structure):
     * rows 2a and 2a+1 share the columns {2b, 2b+1}, so one node id per 2x2 block replaces four column indices
     * and the SpMV streams 9 instead of 12 bytes per non-zero.  Results are bit-identical to the CSR kernels. */
    const int32_t *nb_node_ptr; /* n_rows/2 + 1 (+8 padding): first 2x2 block of node a, absolute (= row_ptr[2a]/4) */
    const int32_t *nb_pcol;     /* nnz/4 (+8 padding): column node of each block, entry k belongs to block nb_node_ptr[0]+k */
    const int16_t *nb_pcol16;   /* or (preferred when non-NULL) nnz/4 (+16 padding): column node - row node, 16 bits     */
    const int32_t *nb_nodeblk;  /* nb_n_blocks + 3 ints */
    int64_t nb_n_blocks;
    int64_t nb_max_pairs;
    /* optional symmetric-lattice storage of exactly `vals` (psfem_spmv_sl_build; NULL / 0 otherwise): for symmetric
     * matrices whose nodes couple with their 3x3 lattice neighbours only (node id = column*sl_m + row, the numbering of
     * Polygones.mesh, Polygones.py:77), the diagonal and the four forward 2x2 blocks per node (19 doubles) stored as a
     * structure of arrays.  The SpMV streams 184 instead of 360 bytes per node and returns the same bits as the CSR
     * kernels.  Used only while sl_src_vals == vals. */
    const double *sl_vals;
    const double *sl_src_vals;  /* the value array sl_vals was built from */
    int64_t sl_m;               /* nodes per lattice column */
    int64_t sl_roff;            /* x node of row node 0 (0, or the halo nodes before the owned rows of a strip) */
    int64_t sl_x_nodes;         /* nodes in the x vector (multiple of sl_m) */
} psfem_csr;

/* ---- mesh: Polygones.mesh, Polygones.py:62-108 ------------------------------------------
 * Nodes of the node columns [col0, col0+ncols) and elements of the cell columns
 * [cell0, cell0+ncells) of the structured mesh mesh(L,l,N,M,offX,offY,etype).  Node ids written
 * to d_conn are LOCAL: global id - col0*m (m = nodes per column), so a strip is a self-contained
 * mesh; col0=0, ncols=n, cell0=0, ncells=N gives the reference's whole mesh.  Coordinates are
 * bit-identical to Polygones.py:77 (product, IEEE division, subtraction; no FMA contraction).
 * d_nodes: ncols*m*2 doubles, d_conn: ncells*M*(2 for triangles)*npe int32.  Either may be NULL. */
int psfem_mesh_rect(double L, double l, int64_t N, int64_t M, double offX, double offY, int etype,
                    int64_t col0, int64_t ncols, int64_t cell0, int64_t ncells,
                    double *d_nodes, int32_t *d_conn, void *stream);

/* ---- symbolic CSR pattern: replaces the dense np.zeros((2n,2n)) of Polygones.py:715,1349 ----
 * Structural pattern of K and M: row 2a+p has the sorted columns {2b+q : node b shares an
 * element with node a}.  Built by a stable radix sort + run-length unique of the npe^2 (a,b)
 * node-pair keys of every element.  The sort also yields the ASSEMBLY PLAN: for every unique pair
 * the contributing (element,i,j) codes in ascending element id -- the order of the reference's
 * scatter loop Polygones.py:717-726 -- which makes the numeric phase a deterministic, atomic-free
 * segmented reduction.
 *   step 1  psfem_pattern_count : sort/unique; d_perm (nel*npe^2 codes e*npe^2+i*npe+j, sorted by
 *                                 pair) is final; returns n_pairs (nnz = 4*n_pairs); state in d_ws
 *   step 2  psfem_pattern_fill  : writes row_ptr, col_idx and the rest of the plan (same d_ws) */
int psfem_pattern_workspace(int64_t nel, int npe, int64_t n_nodes, size_t *bytes);
int psfem_pattern_count(const int32_t *d_conn, int64_t nel, int npe, int64_t n_nodes, int32_t *d_perm,
                        void *d_ws, size_t ws_bytes, int64_t *h_n_pairs, void *stream);
int psfem_pattern_fill(int64_t nel, int npe, int64_t n_nodes, int64_t n_pairs, void *d_ws, size_t ws_bytes,
                       int32_t *d_row_ptr,  /* 2*n_nodes+1 */
                       int32_t *d_col_idx,  /* 4*n_pairs   */
                       int32_t *d_node_ptr, /* n_nodes+1 : first pair of each node row */
                       int32_t *d_pair_row, /* n_pairs   : row node of each pair        */
                       int32_t *d_pair_ptr, /* n_pairs+1 : segment bounds into d_perm   */
                       void *stream);

/* ---- numeric assembly: Polygones.global_stiffness_matrix :710-728 / global_mass_matrix :1344-1364
 * what: bit0 = K (element_stiffness_matrix :566-655, T3 :875-883), bit1 = M (element_mass_matrix
 * :1250-1342, T3 :1228-1235).  d_K_vals / d_M_vals: nnz doubles on the pattern above (M keeps
 * explicit zeros at u-v couplings).  h_bad_element: first element (lowest id) with |detJ|<1e-10,
 * else -1; the call then returns PSFEM_EJACOBIAN.  Synchronises the stream. */
int psfem_assemble_workspace(int64_t nel, int etype, int what, size_t *bytes);
int psfem_assemble(const double *d_nodes, const int32_t *d_conn, int64_t nel, int etype, int64_t n_nodes,
                   double E, double nu, double rho, double t, int what, int64_t n_pairs,
                   const int32_t *d_node_ptr, const int32_t *d_pair_row, const int32_t *d_pair_ptr,
                   const int32_t *d_perm, double *d_K_vals, double *d_M_vals,
                   void *d_ws, size_t ws_bytes, int64_t *h_bad_element, void *stream);

/* ---- tile-fused numeric assembly for T3 / Q4 (the fast path of psfem_assemble) ------------------
 * A CTA owns a tile of tile_nodes_R (even, 32..256) node rows that are neighbours in space (Morton order
 * of the coordinates, so any mesh tiles well).  Every element touching the tile is computed once, in
 * registers; the tile's block of CSR values lives in shared memory and the elements add their matrix rows
 * into it in rounds (round r of a node = its r-th incident element in ascending id: all rows of a round
 * belong to distinct nodes, so it is race-free without atomics, and every entry is summed in ascending
 * element id like Polygones.py:717-726 => bit-reproducible); the finished rows are written to HBM once.  The kernel is persistent and software pipelined (cp.async coordinate gathers + TMA bulk
 * copies of the plan, one tile ahead).
 * psfem_tileplan_build derives the plan from the coordinates, the connectivity and the CSR pattern
 * (d_node_ptr of psfem_pattern_fill, d_col_idx):
 *   d_tile_nodes    n_nodes       node ids sorted by (tile, id); tile t = entries [t*R, (t+1)*R)
 *   d_tile_elem_ptr n_tiles+1     n_tiles = ceil(n_nodes / R)
 *   d_tile_elems    nel*npe       (upper bound) element ids per tile, ascending; h_info[0] used
 *   d_tile_hdr      8*n_tiles     {e0, ne, nn, npairs, nrounds, 0,0,0}
 *   d_tile_conn     4*nel*npe     (upper bound) int4 connectivity of every tile element; bits 28..31 of
 *                                 entry i = scatter round of row i (15 = row node not owned by the tile)
 *   d_tile_rec      32 B*nel*npe  (upper bound) scatter record of every tile element: double2 offsets of its
 *                                 rows inside the tile's value block (kept in shared memory in the global layout)
 *   d_nmeta         2*n_tiles*R   {first CSR pair of the node, pair prefix inside its tile}
 *   h_info[6]       {tile elements used, max elements, max pairs, max rounds per tile, usable (1/0), mean elements}
 * Same summation order as psfem_assemble (ascending element id).
 * Work space of psfem_assemble_tiled: 16 bytes. */
int psfem_tileplan_workspace(int64_t nel, int npe, int64_t n_nodes, int tile_nodes_R, size_t *bytes);
int psfem_tileplan_build(const double *d_nodes, const int32_t *d_conn, int64_t nel, int npe, int64_t n_nodes,
                         int tile_nodes_R, const int32_t *d_node_ptr, const int32_t *d_col_idx,
                         int32_t *d_tile_nodes, int32_t *d_tile_elem_ptr, int32_t *d_tile_elems, int32_t *d_tile_hdr,
                         int32_t *d_tile_conn, void *d_tile_rec, int32_t *d_nmeta, int64_t *h_info,
                         void *d_ws, size_t ws_bytes, void *stream);
int psfem_assemble_tiled(const double *d_nodes, int64_t nel, int etype, int64_t n_nodes,
                         double E, double nu, double rho, double t, int what, int tile_nodes_R,
                         const int64_t *h_plan_info, const int32_t *d_tile_hdr, const int32_t *d_tile_conn,
                         const void *d_tile_rec, const int32_t *d_tile_elems, const int32_t *d_nmeta,
                         double *d_K_vals, double *d_M_vals, void *d_ws, size_t ws_bytes,
                         int64_t *h_bad_element, void *stream);

/* ---- lattice assembly: the fast path for meshes with the topology of Polygones.mesh (Polygones.py:62-108) ----
 * psfem_lattice_detect: h_nm = {n, m} (node columns, nodes per column) when EVERY element equals the formula of
 * Polygones.py:83-85 (T3) / :100 (Q4) for node id = i*m+j, else {0, 0}.  Coordinates are not looked at: a
 * jittered or otherwise deformed structured mesh is still a lattice.  d_ws: >= 16 bytes.  Synchronises.
 * psfem_assemble_lattice (T3, Q4): same result as psfem_assemble / psfem_assemble_tiled (values summed in ascending
 * element id, Polygones.py:717-726) on the pattern of psfem_pattern_fill, but plan-free: a warp marches along i
 * over a strip of 31 node rows, element matrices and the open node column live in registers, the rows owned by the
 * neighbour lane travel by shuffle, finished node columns leave by bulk async stores.  d_ws: >= 32 bytes. */
int psfem_lattice_detect(const int32_t *d_conn, int64_t nel, int etype, int64_t n_nodes, int64_t *h_nm /* [2] */,
                         void *d_ws, size_t ws_bytes, void *stream);
/* Symbolic phase of a lattice in closed form (no sort): the same row_ptr / col_idx / node_ptr that
 * psfem_pattern_count + psfem_pattern_fill produce for that connectivity (tests compare them bit for bit), written
 * by one kernel at HBM speed.  d_row_ptr == NULL: only returns n_pairs (nnz = 4*n_pairs) for sizing. */
int psfem_lattice_pattern(int64_t n, int64_t m, int etype, int32_t *d_row_ptr /* 2nm+1 */, int32_t *d_col_idx /* 4*n_pairs */,
                          int32_t *d_node_ptr /* nm+1 */, int64_t *h_n_pairs, void *stream);
int psfem_assemble_lattice(const double *d_nodes, int64_t n, int64_t m, int etype,
                           double E, double nu, double rho, double t, int what,
                           double *d_K_vals, double *d_M_vals, void *d_ws, size_t ws_bytes,
                           int64_t *h_bad_element, void *stream);

/* The same numeric assembly (Q4 lattices: global_stiffness_matrix :710-728 / global_mass_matrix :1344-1364) written STRAIGHT
 * INTO the symmetric-lattice operand of the solvers (psfem_spmv_sl_build's layout: 19 doubles per node): for flows that
 * only solve, K never exists as CSR -- no value array, no column indices, no reduced copy (2.55 instead of ~12 GB and ~15 ms
 * of set-up at the 4096 x 4096 plate; the 16384 x 4096 strip fits ONE GPU).  Stores the node columns [c_lo, c_lo + S);
 * c_lo = 1 is the reduction by a clamped first column.  d_sl_vals: S * 19 * ((m+3)&~3) doubles.  what: 1 = K, 2 = M.
 * Bit-identical to psfem_assemble_lattice + psfem_spmv_sl_build.  d_ws >= 32 bytes.  Synchronises. */
int psfem_assemble_lattice_sl(const double *d_nodes, int64_t n, int64_t m, int etype, double E, double nu, double rho, double t,
                              int what, int64_t c_lo, int64_t S, double *d_sl_vals, size_t bytes, void *d_ws, size_t ws_bytes,
                              int64_t *h_bad_element, void *stream);

/* element matrices only: Polygones.element_stiffness_matrix :566, element_mass_matrix :1250.
 * d_Ke / d_Me: nel*(2npe)^2 doubles, one row-major (2npe x 2npe) matrix per element.
 * Work space: psfem_assemble_workspace(nel, etype, what). */
int psfem_element_matrices(const double *d_nodes, const int32_t *d_conn, int64_t nel, int etype,
                           double E, double nu, double rho, double t, int what, double *d_Ke, double *d_Me,
                           void *d_ws, size_t ws_bytes, int64_t *h_bad_element, void *stream);

/* compute_B_matrix_at_point, Polygones.py:657-708: strain-displacement matrix B (3 x 2npe, rows eps_x, eps_y, gamma_xy),
 * Jacobian J and det J of ONE element (d_conn1: its npe node ids) at the reference point (xi, eta); shape functions of
 * Polygones.py:441-518 for all four element kinds.  d_out / h_out: 6*npe + 6 doubles = B row-major, J00 J01 J10 J11,
 * detJ, flag.  Returns PSFEM_EJACOBIAN when |detJ| < 1e-10 (:686-687).  Synchronises. */
int psfem_b_matrix_at_point(const double *d_nodes, const int32_t *d_conn1, int etype, double xi, double eta,
                            double *d_out, double *h_out, void *stream);

/* ---- boundary conditions: Polygones.apply_boundary_conditions :989-1006, reduce :898-911 ----
 * psfem_free_map: d_constrained (nc DOF ids, duplicates and out-of-range ids allowed, as the list
 * membership test of :1000 allows them) -> d_new_id[n] (reduced index or -1) and d_free_dofs
 * (ascending complement); returns n_free.
 * psfem_reduce_count/fill: K[np.ix_(free,free)] on CSR; d_keep[k] = position in the full value
 * array of reduced entry k, so K, M and K_eff values reduce with psfem_gather.
 * Work space for all three: psfem_scan_workspace(n). */
int psfem_scan_workspace(int64_t n, size_t *bytes);
int psfem_free_map(int64_t n, const int32_t *d_constrained, int64_t nc, int32_t *d_new_id,
                   int32_t *d_free_dofs, int64_t *h_n_free, void *d_ws, size_t ws_bytes, void *stream);
int psfem_reduce_count(int64_t n, const int32_t *d_row_ptr, const int32_t *d_col_idx, const int32_t *d_new_id,
                       int64_t n_free, int32_t *d_row_ptr_red, int64_t *h_nnz_red,
                       void *d_ws, size_t ws_bytes, void *stream);
int psfem_reduce_fill(int64_t n, const int32_t *d_row_ptr, const int32_t *d_col_idx, const int32_t *d_new_id,
                      const int32_t *d_row_ptr_red, int32_t *d_col_idx_red, int32_t *d_keep, void *stream);
/* sanity of a CSR operand that arrives from the host (replaces scipy's host-side has_sorted_indices scan):
 * *h_flags bit0 = columns not strictly ascending within a row, bit1 = column out of [0, n_cols), bit2 = row_ptr not
 * monotone.  d_ws: >= 16 bytes.  Synchronises. */
int psfem_csr_check(int64_t n, int64_t n_cols, const int32_t *d_row_ptr, const int32_t *d_col_idx, int64_t nnz,
                    int64_t *h_flags, void *d_ws, size_t ws_bytes, void *stream);
int psfem_gather(const double *d_src, const int32_t *d_index, int64_t n, double *d_dst, void *stream);
/* reconstruct_full_vector(s) Polygones.py:913-936: d_full[v, free[i]] = d_red[v, i], zeros elsewhere */
int psfem_scatter_free(const double *d_red, const int32_t *d_free_dofs, int64_t n_free, int64_t n_full,
                       int64_t nvec, double *d_full, void *stream);

/* ---- sparse kernels used by the static / Newmark / modal drivers -------------------------
 * SpMV stages row blocks of the CSR stream through shared memory (coalesced vals/col_idx reads,
 * per-row reduction from shared memory in column order). */
int psfem_spmv_plan_size(int64_t n, int64_t nnz, int64_t *n_blocks);
int psfem_spmv_plan(int64_t n, const int32_t *d_row_ptr, int64_t nnz, int32_t *d_rowblk /* n_blocks+2 */,
                    int64_t n_blocks, int64_t *h_max_row_nnz, void *stream);
/* node-block view of a CSR operand: h_info[0] = 1 if the matrix has the structure (else the outputs are unused),
 * h_info[1] = largest number of 2x2 blocks in a node row, h_info[2] = largest |column node - row node|.
 * d_node_ptr: n/2+9 ints, d_pcol: nnz/4+8 ints, d_nodeblk: n_blocks+4 ints with n_blocks from
 * psfem_spmv_nb_plan_size.  Synchronises.  When h_info[2] < 32768, psfem_spmv_nb_compress turns d_pcol into 16-bit
 * offsets (d_pcol16: nnz/4+16 int16): 8.5 instead of 9 bytes per non-zero. */
int psfem_spmv_nb_plan_size(int64_t n, int64_t nnz, int64_t *n_blocks);
int psfem_spmv_nb_plan(int64_t n, const int32_t *d_row_ptr, const int32_t *d_col_idx, int64_t nnz,
                       int32_t *d_node_ptr, int32_t *d_pcol, int32_t *d_nodeblk, int64_t n_blocks,
                       int64_t *h_info /* [3] */, void *stream);
int psfem_spmv_nb_compress(int64_t n, const int32_t *d_node_ptr, const int32_t *d_pcol, int64_t nnz,
                           int16_t *d_pcol16, void *stream);
/* y = alpha*A*x + gamma*z   (z may be NULL; y may alias z) */
/* Symmetric-lattice storage for the SpMV of PCG (replaces the same dense products as psfem_spmv: main.py:310,383-384
 * and the solves of Statictest.py:69, main.py:286,314 through PCG).
 *   psfem_spmv_sl_candidates: proposals for the lattice column height m of a node-block operand (h_m[8], *h_count of them;
 *                             0 when the operand cannot be a lattice); synchronises
 *   psfem_spmv_sl_size:       bytes of the storage for (row nodes, row_node_offset, m); allocate 256 bytes more -- and, for
 *                             the TMA-staged kernels (psfem_cg_sl_*, the SpMV itself), keep 256 readable bytes IN FRONT of
 *                             d_sl_vals and 2 KB behind the storage: their 116-row window copies start two rows before a
 *                             strip group and run past the last one (the Python side allocates that slack)
 *   psfem_spmv_sl_build:      checks that A is a 3x3-lattice matrix with column height m whose 2x2 blocks (a,b) and (b,a)
 *                             are transposes (tol = 0: bit for bit, which makes the SpMV bit-identical to the CSR kernels;
 *                             tol > 0: within tol * the node's diagonal) and writes the storage; *h_ok = 1 on success.
 * row_node_offset = x node of the first row node (rows of a strip start after its left halo column), x_nodes = nodes
 * in the operand vector.  The caller then sets sl_vals / sl_src_vals / sl_m / sl_roff / sl_x_nodes in psfem_csr. */
int psfem_spmv_sl_candidates(const psfem_csr *A, int64_t row_node_offset, int64_t x_nodes, int64_t *h_m, int64_t *h_count,
                             void *stream);
int psfem_spmv_sl_size(int64_t n_row_nodes, int64_t row_node_offset, int64_t m, size_t *bytes);
int psfem_spmv_sl_build(const psfem_csr *A, int64_t row_node_offset, int64_t x_nodes, int64_t m, double tol,
                        double *d_sl_vals, size_t bytes, int64_t *h_ok, void *stream);
int psfem_spmv(const psfem_csr *A, const double *d_x, double alpha, double gamma, const double *d_z,
               double *d_y, void *stream);
/* y = alpha*A1*x1 + beta*A2*x2 + gamma*z, A2 = (pattern of A, d_vals2): Newmark RHS main.py:310 */
int psfem_spmv2(const psfem_csr *A, const double *d_vals2, const double *d_x1, const double *d_x2,
                double alpha, double beta, double gamma, const double *d_z, double *d_y, void *stream);
/* diag[r] = A[r, r+col_offset] (invert != 0: its reciprocal = the Jacobi preconditioner) */
int psfem_extract_diag(const psfem_csr *A, int64_t col_offset, int invert, double *d_diag, void *stream);

/* Jacobi-preconditioned CG: replaces np.linalg.solve / scipy.linalg.solve / lu_solve of
 * Statictest.py:69, StaticTest2.py:71, Beamexemple.py:63, main.py:286,314.
 * x0 = d_x on entry; stops when ||r||_2 <= rtol*||b||_2 (decided on the device, so the iteration
 * count is exact); the host polls every check_every iterations.  Systems of <= 16384 rows run in a
 * single-CTA kernel (one launch per solve).  h_info: [0]=iterations, [1]=relative residual, [2]=||b||. */
int psfem_pcg_workspace(int64_t n, int64_t n_blocks, size_t *bytes);
int psfem_pcg_jacobi(const psfem_csr *A, const double *d_b, double *d_x, double rtol, int64_t maxit,
                     int64_t check_every, double *h_info, void *d_ws, size_t ws_bytes, void *stream);

/* The same solve on the UNREDUCED matrix: DOFs with d_new_id[i] < 0 (psfem_free_map) stay in the system as inert rows
 * (D^-1 = 0, residual, direction and right-hand side pinned to zero), so the iterates on the free DOFs are those of the
 * reduced system K[np.ix_(free,free)] (Polygones.py:1003) while the matrix keeps its 2x2 node-block structure for the
 * 9-byte SpMV whatever the boundary conditions (a roller constrains one DOF of a node), and no reduced copy of the
 * matrix is built.  d_b: full-length load (entries at constrained DOFs ignored), d_x: full-length start / solution
 * (zeros at constrained DOFs).  For systems above 16384 rows. */
int psfem_pcg_jacobi_masked(const psfem_csr *A, const double *d_b, double *d_x, const int32_t *d_new_id, double rtol,
                            int64_t maxit, int64_t check_every, double *h_info, void *d_ws, size_t ws_bytes, void *stream);

/* ---- direct static solve of SMALL banded systems + iterative refinement on a compensated residual --------------
 * The reference's static solves are dense LAPACK LU (np.linalg.solve: TestBeam/Statictest.py:69, StaticTest2.py:71;
 * scipy.linalg.solve: Beamexemple.py:63; solve_banded: 2Dexemple.py:36).  For systems of up to
 * psfem_direct_limits() rows / half bandwidth this entry does the same kind of solve on the device: banded Cholesky
 * (node id = column*m + row makes K banded), then `refine_rounds` rounds of  r = b - K x  (accumulated in
 * double-double),  K d = r,  x += d.  x ends within a few ulp of the exact solution of the fp64 system, so its distance
 * to the reference's answer is the reference's own LU rounding error (4.7e-11 on the README cantilever, kappa = 7e7).
 *   psfem_band_width:   max |row - col| over the pattern (d_ws >= 16 bytes; synchronises)
 *   psfem_direct_solve: h_info[0] = refinement rounds done, [1] = ||b - K x|| / ||b|| (compensated), [2] = ||b||,
 *                       [3] = status, [4] = ||last correction|| / ||x||.  Returns PSFEM_ENOCONV when a Cholesky pivot
 *                       is not positive.  Synchronises.
 *   psfem_residual_dd:  r = b - A x with the same compensated accumulation for ANY CSR operand: the residual of the
 *                       PCG-based refinement of systems beyond the band solver's limits. */
int psfem_band_width(const psfem_csr *A, int64_t *h_half_bandwidth, void *d_ws, size_t ws_bytes, void *stream);
int psfem_direct_limits(int64_t *max_rows, int64_t *max_half_bandwidth);
int psfem_direct_workspace(int64_t n, int64_t half_bandwidth, size_t *bytes);
int psfem_direct_solve(const psfem_csr *A, const double *d_b, double *d_x, int64_t half_bandwidth, int refine_rounds,
                       double *h_info /* [5] */, void *d_ws, size_t ws_bytes, void *stream);
/* the factorisation alone (it stays in d_ws) and solves with it: K is factored ONCE for the shift-invert Lanczos of the modal
 * driver (freeModes.py:74 / Beamexemple.py:109 do the same through LAPACK / ARPACK) */
int psfem_direct_factor(const psfem_csr *A, int64_t half_bandwidth, void *d_ws, size_t ws_bytes, void *stream);
int psfem_direct_substitute(const psfem_csr *A, const double *d_b, double *d_x, int64_t half_bandwidth, int refine_rounds,
                            double *h_info /* [5] */, void *d_ws, size_t ws_bytes, void *stream);
int psfem_residual_dd(const psfem_csr *A, const double *d_x, const double *d_b, double *d_r, void *stream);
/* measured fp64 FMA throughput of the current device in TFLOP/s (a short micro-kernel on all SMs; d_ws >= 16 bytes): the
 * assembly kernels (Polygones.py:566-708 element arithmetic) are co-bound by the fp64 pipe, so bench.py reports them against
 * the HBM roofline AND against this peak.  Synchronises. */
int psfem_fp64_peak(double *h_tflops, void *d_ws, size_t ws_bytes, void *stream);

/* np.allclose(A, A.T, rtol, atol) on CSR -- the symmetry test of main.py:65-66, Beamexemple.py:48, SparseTest.py:37 without
 * a dense transpose: every stored entry is compared with its mirror entry (0 when not stored).  *h_violations = entries
 * with |A[r,c] - A[c,r]| > atol + rtol |A[c,r]|, *h_max_abs_diff = largest |A[r,c] - A[c,r]|.  d_ws >= 16 bytes.
 * Needs sorted column indices.  Synchronises. */
int psfem_csr_symmetry(const psfem_csr *A, double rtol, double atol, int64_t *h_violations, double *h_max_abs_diff,
                       void *d_ws, size_t ws_bytes, void *stream);

/* ---- Jacobi-PCG with ONE KERNEL PER ITERATION on symmetric-lattice operands ------------------------------------------
 * Same solves as psfem_pcg_jacobi / psfem_pcg_jacobi_masked (Statictest.py:69, main.py:286,314) for operands that carry
 * their symmetric-lattice storage (psfem_spmv_sl_build: the matrices of Polygones.mesh meshes), in the Chronopoulos-Gear
 * form of the recurrence (one reduction per iteration; same Krylov iterates in exact arithmetic):
 *     p = u + beta p;  s = w + beta s;  x += alpha p;  r -= alpha s;  u = D^-1 r;  w = A u;  {r.u, w.u, r.r} -> alpha', beta'
 * The pointwise part rides on the marching warps of the SpMV, operands staged through a TMA + mbarrier shared-memory ring:
 * 296 B of HBM traffic per node and iteration instead of 344 B in three kernels, D^-1 taken from the streamed diagonal.  d_new_id != NULL: constrained DOFs stay as masked rows.
 *   psfem_cg_sl_usable     *h_usable = 1 when A can take this path (symmetric-lattice storage of A->vals, rows = columns)
 *   psfem_cg_sl_start      x0 = d_x on entry; r0, u0, w0, alpha0 (uses the SpMV kernel twice)
 *   psfem_cg_sl_iterations k launches; iters_done = iterations launched since the start (selects the buffer parity)
 *   psfem_cg_sl_state      h_info = {iterations, relative residual, ||b||, done, status}; synchronises
 *   psfem_cg_sl_finish     x is written every other iteration only (x += alpha_{i-1} p_{i-1} + alpha_i p_i): this adds the
 *                          pending half, if any; call it before reading d_x after psfem_cg_sl_iterations
 *   psfem_cg_sl_solve      start + batches of check_every iterations until ||r|| <= rtol ||b|| (decided on the device);
 *                          h_info = {iterations, relative residual of the recurrence, ||b||}
 *   psfem_cg_sl_solve_checked  the same, then the TRUE residual b - A x (scipy.sparse.linalg.cg, the solver of
 *                          Statictest.py:69, stops on its recurrence as well; over 10^4..10^5 iterations the recurred residual
 *                          drifts below the true one): the solve is started again from x and continued until the true
 *                          residual passes the test, stops improving by a factor 2 per cycle, or `polish` cycles are spent.
 *                          polish = 0: evaluate only; polish < 0: none.  h_info = {iterations of all cycles, recurrence relres
 *                          at the first convergence, ||b||, true relres at the last evaluation (-1: none), cycles continued,
 *                          iterations spent in them} */
int psfem_cg_sl_usable(const psfem_csr *A, int *h_usable);
int psfem_cg_sl_workspace(int64_t n, size_t *bytes);
int psfem_cg_sl_start(const psfem_csr *A, const double *d_b, double *d_x, const int32_t *d_new_id, double rtol,
                      void *d_ws, size_t ws_bytes, void *stream);
int psfem_cg_sl_iterations(const psfem_csr *A, double *d_x, int masked, int64_t k, int64_t iters_done, void *d_ws,
                           size_t ws_bytes, void *stream);
int psfem_cg_sl_state(void *d_ws, int64_t n, double *h_info /* [5] */, void *stream);
int psfem_cg_sl_finish(double *d_x, int64_t n, void *d_ws, void *stream);
int psfem_cg_sl_solve(const psfem_csr *A, const double *d_b, double *d_x, const int32_t *d_new_id, double rtol, int64_t maxit,
                      int64_t check_every, double *h_info /* [3] */, void *d_ws, size_t ws_bytes, void *stream);
int psfem_cg_sl_solve_checked(const psfem_csr *A, const double *d_b, double *d_x, const int32_t *d_new_id, double rtol,
                              int64_t maxit, int64_t check_every, int64_t polish, double *h_info /* [6] */, void *d_ws,
                              size_t ws_bytes, void *stream);

/* The kernels of one PCG iteration, exposed so the strip-partitioned multi-GPU driver can put its
 * halo exchange and all-reduces between them (scalars stay on the device):
 *   init    : p = dinv*r ; out3 = {r.dinv r, r.r, b.b}  (local sums)
 *   spmv_dot: q = A p_ext ; *d_pq = p_own . q  (p_own = p_ext + x_offset)
 *   update  : alpha = *d_rz / *d_pq ; r -= alpha q ; out2 = {r.dinv r, r.r}
 *   pupdate : alpha = *d_rz_old / *d_pq ; beta = *d_rz_new / *d_rz_old ; x += alpha p ; p = dinv*r + beta p
 *             (the x step lives here because this kernel streams p anyway: 80 instead of 88 bytes per DOF and iteration)
 * d_partials: max(n_blocks, 1184)*3 doubles, d_counter: zero-initialised uint32. */
int psfem_pcg_init(int64_t n, const double *d_r, const double *d_dinv, const double *d_b, double *d_p,
                   double *d_out3, double *d_partials, uint32_t *d_counter, void *stream);
int psfem_pcg_spmv_dot(const psfem_csr *A, const double *d_p_ext, int64_t x_offset, double *d_q, double *d_pq,
                       double *d_partials, uint32_t *d_counter, void *stream);
int psfem_pcg_update(int64_t n, double *d_r, const double *d_q,
                     const double *d_dinv, const double *d_rz, const double *d_pq, double *d_out2,
                     double *d_partials, uint32_t *d_counter, void *stream);
/* psfem_pcg_update for systems that keep constrained DOFs as masked rows (d_dinv = 0 there): their residual stays 0 */
int psfem_pcg_update_masked(int64_t n, double *d_r, const double *d_q,
                            const double *d_dinv, const double *d_rz, const double *d_pq, double *d_out2,
                            double *d_partials, uint32_t *d_counter, void *stream);
int psfem_pcg_pupdate(int64_t n, double *d_p, double *d_x, const double *d_r, const double *d_dinv,
                      const double *d_rz_old, const double *d_rz_new, const double *d_pq, void *stream);

/* ---- strip-partitioned Jacobi-PCG with the collectives FUSED into its kernels (multi-GPU, one process per GPU) ----
 * Replaces, per iteration, two NCCL all-reduces and one halo send/recv by stores over NVLink/NVSwitch peer memory:
 *   - every rank owns a small "board" and its extended p vector in memory that the other ranks map
 *     (psfem_peer_alloc -> 64-byte CUDA IPC handle, exchanged by the host through torch.distributed;
 *     psfem_peer_open on the other ranks);
 *   - the SpMV kernel publishes its p.q partial to every board, the update kernel spins on its OWN board, adds the
 *     partials in rank order (bit-identical on every rank) and publishes {r.z, r.r}; the p-update kernel reads them,
 *     computes p and stores the first / last owned node column straight into the neighbours' extended vectors,
 *     then raises their halo flag, which the next SpMV waits for (inside the two edge column segments of the
 *     symmetric-lattice kernel; at the end of the p-update kernel for the other SpMV kernels).
 *   Scalars travel as 8-byte words {32 data bits | 32-bit sequence number}: no fence between data and flag.
 * Sequence numbers only grow: psfem_dist_pcg_start takes seq >= 1, the following iterations seq+1, seq+2, ...
 * (the host keeps the counter; it is never reset while the boards live).  Same recurrence and the same per-rank
 * kernels as psfem_pcg_spmv_dot / _update / _pupdate. */
#define PSFEM_MAX_RANKS 8
typedef struct psfem_dist_ctx {
    int32_t rank, world, has_left, has_right;
    void *board[PSFEM_MAX_RANKS];         /* board[w] of rank w (peer mapped); board[rank] is this rank's own        */
    double *p_ext;                         /* this rank's extended p vector (peer-visible allocation)                */
    double *left_p_ext, *right_p_ext;      /* the neighbours' extended p vectors (peer mapped) or NULL                */
    int64_t own_lo, n_own;                 /* owned entries of p_ext: [own_lo, own_lo + n_own)                        */
    int64_t send_left_n, left_dst_off;     /* first send_left_n owned entries -> left_p_ext[left_dst_off ...]         */
    int64_t send_right_n, right_dst_off;   /* last send_right_n owned entries -> right_p_ext[right_dst_off ...]       */
    double *x, *r, *q, *dinv;              /* owned rows                                                              */
    double *out3;                          /* 8 doubles of local device memory: global scalars of the recurrence      */
    double *partials;                      /* max(n_blocks, 1184)*3 doubles                                           */
    uint32_t *counter;                     /* 4 zero-initialised uint32                                               */
    int32_t masked;                        /* 1: constrained DOFs stay in the strip matrix as masked rows (dinv = 0 there) */
} psfem_dist_ctx;
int psfem_peer_board_bytes(size_t *bytes);
int psfem_peer_alloc(size_t bytes, void **d_ptr, unsigned char *handle64);   /* zero-filled */
int psfem_peer_open(const unsigned char *handle64, void **d_ptr);
int psfem_peer_close(void *d_ptr);
int psfem_peer_free(void *d_ptr);
/* x = 0, r = b, p = D^-1 r (halo pushed), {r.z, r.r, b.b} published with sequence number seq */
int psfem_dist_pcg_start(const psfem_dist_ctx *ctx, const double *d_b, uint64_t seq, void *stream);
/* k iterations with sequence numbers seq_first .. seq_first+k-1 (seq_first = previous sequence number + 1) */
int psfem_dist_pcg_iterations(const psfem_csr *A, const psfem_dist_ctx *ctx, int64_t k, uint64_t seq_first, void *stream);
/* diagnostic twin of psfem_dist_pcg_iterations: mean milliseconds of {SpMV+dot, update, p-update} (CUDA events
 * around every kernel; synchronises) */
int psfem_dist_pcg_profile(const psfem_csr *A, const psfem_dist_ctx *ctx, int64_t k, uint64_t seq_first, double *h_ms3, void *stream);
/* global {r.z, r.r, b.b} of the step with sequence number seq (b.b is valid for the start step only); synchronises */
int psfem_dist_pcg_residual(const psfem_dist_ctx *ctx, uint64_t seq, double *h_out3, void *stream);

/* ---- the fused one-kernel PCG iteration (psfem_cg_sl_*) on a strip of the multi-GPU partition --------------------------
 * One launch per iteration and ONE cross-GPU scalar exchange ({r.u, w.u, r.r} through the peer boards, LL words); the halo
 * is u of the first / last owned lattice column, computed first by the edge column segments and stored into the
 * neighbour's LL halo buffer ({32 data bits | 32-bit sequence number} words polled by exactly the lanes that need them:
 * no fence, no flag).  Replaces psfem_dist_pcg_* (three launches, two exchanges) when the strip operand carries its
 * symmetric-lattice storage.  State lives in a psfem_cg_sl_workspace block; psfem_cg_sl_finish adds the pending x half.
 *   psfem_dist_cg_halo_bytes  size of a rank's LL halo buffer (psfem_peer_alloc, zero-filled) for column height m
 *   psfem_dist_cg_start       x = 0, r = b and the start pass (launch 0, sequence number seq >= 1); masked strips pass the free
 *                             map of their owned rows (d_new_id[i] < 0: constrained DOF kept in the matrix)
 *   psfem_dist_cg_iterations  k launches; launches_done counts launches since the start (>= 1), seq_first = seq + launches_done
 *   psfem_dist_cg_state       {iterations, relative residual, ||b||, done, status}, global; reports a peer time-out
 *   psfem_dist_cg_apply       y = A x over the strips (owned rows in, owned rows out; one launch of the same kernel in apply
 *                             mode, sequence number seq): the TRUE residual b - A x of a finished strip solve.  Overwrites
 *                             the recurrence vectors (continue with psfem_dist_cg_start on the residual), keeps the state */
typedef struct psfem_dist_cg_ctx {
    int32_t rank, world, has_left, has_right;
    void *board[PSFEM_MAX_RANKS];          /* psfem_peer_board_bytes boards, board[rank] is this rank's own                */
    void *ll_mine;                         /* this rank's LL halo buffer                                                  */
    void *ll_left, *ll_right;              /* the neighbours' LL halo buffers (peer mapped) or NULL                        */
    int64_t m;                             /* nodes per lattice column                                                    */
    double *x;                             /* owned rows of the solution                                                  */
    uint32_t *counter;                     /* 4 zero-initialised uint32; [3] = peer time-out flag                          */
    int32_t masked;
} psfem_dist_cg_ctx;
int psfem_dist_cg_halo_bytes(int64_t m, size_t *bytes);
int psfem_dist_cg_start(const psfem_csr *A, const psfem_dist_cg_ctx *ctx, const double *d_b, const int32_t *d_new_id,
                        double rtol, uint64_t seq, void *d_ws, size_t ws_bytes, void *stream);
int psfem_dist_cg_iterations(const psfem_csr *A, const psfem_dist_cg_ctx *ctx, int64_t k, int64_t launches_done,
                             uint64_t seq_first, void *d_ws, size_t ws_bytes, void *stream);
int psfem_dist_cg_state(const psfem_dist_cg_ctx *ctx, int64_t n, void *d_ws, double *h_info /* [5] */, void *stream);
int psfem_dist_cg_apply(const psfem_csr *A, const psfem_dist_cg_ctx *ctx, const double *d_x, double *d_y, uint64_t seq,
                        void *d_ws, size_t ws_bytes, void *stream);

/* ---- BLAS-1 helpers for the drivers (Lanczos, K_eff = K + c M, energies) ------------------ */
int psfem_axpby(int64_t n, double a, const double *d_x, double b, const double *d_y /* may be NULL */,
                double *d_out, void *stream);
int psfem_dot_workspace(int64_t n, size_t *bytes);
int psfem_dot(int64_t n, const double *d_x, const double *d_y, double *h_out, void *d_ws, size_t ws_bytes, void *stream);
/* out[j] = V[j,:] . w  for j<m  (V row-major m x n): Gram-Schmidt coefficients */
int psfem_multi_dot(int64_t m, int64_t n, const double *d_V, const double *d_w, double *d_out, void *stream);
/* w -= sum_j c[j] V[j,:] */
int psfem_multi_axpy(int64_t m, int64_t n, const double *d_V, const double *d_c, double *d_w, void *stream);

/* ---- Newmark-beta as implemented by main.py:264-332 (NOT the textbook scheme) -------------
 * Runs steps 1..n_steps-1 on the reduced system: K, M and Keff = K + M/(beta dt^2) are three operands on ONE
 * pattern (same row_ptr / col_idx / plans, their own vals and, optionally, their own symmetric-lattice storage).
 * Load of step i: d_fshape * h_fscale[i] (main.py:226-251 is a fixed nodal pattern times a ramp).
 * d_U, d_V, d_A: column c (n doubles each) holds step c*store_every; any may be NULL.
 * d_energy: n_steps x 3 (kinetic, potential, total; main.py:377-385) or NULL.
 * Solves: Jacobi-PCG to rtol, warm-started from the previous acceleration.
 * h_info: [0]=total PCG iterations, [1]=max relres, [2]=stabiliser triggers (main.py:324-332). */
int psfem_newmark_workspace(int64_t n, int64_t n_blocks, size_t *bytes);
int psfem_newmark_run(const psfem_csr *K, const psfem_csr *M, const psfem_csr *Keff,
                      const double *d_fshape, const double *h_fscale, int64_t n_steps,
                      double dt, double gamma, double beta, double alpha_R, double beta_R,
                      const double *d_a0, int64_t store_every,
                      double *d_U, double *d_V, double *d_A, double *d_energy,
                      double rtol, int64_t maxit, double *h_info,
                      void *d_ws, size_t ws_bytes, void *stream);

/* ---- modified explicit Euler as implemented by main.py:334-372 (the `use_newmark = False` branch) ----
 * u_i = u_{i-1} + dt v_{i-1};  a_i = M^-1 (F_i - K u_i);  v_i = v_{i-1} + dt/2 (a_{i-1} + a_i);  stabiliser with
 * threshold 10 (main.py:365-370).  Same buffers, outputs and h_info as psfem_newmark_run; the solves are with the mass
 * matrix (Jacobi-PCG, or the banded Cholesky factor of M for small banded systems).  The reference steps with
 * times[i] - times[i-1] of a linspace; this entry uses the constant dt. */
int psfem_explicit_run(const psfem_csr *K, const psfem_csr *M, const double *d_fshape, const double *h_fscale,
                       int64_t n_steps, double dt, const double *d_a0, int64_t store_every,
                       double *d_U, double *d_V, double *d_A, double *d_energy,
                       double rtol, int64_t maxit, double *h_info, void *d_ws, size_t ws_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PSFEM_H */
