// K1-K6: P1/P2 tetrahedral assembly of A, B, C, the flame vectors, the CSR pattern
// and the point location -- what FFCx tabulate_tensor + DOLFINx assemble_matrix /
// assemble_vector / SparsityPattern / determine_point_ownership do for the reference
// (helmholtz_x/acoustic_matrices.py:101-123, helmholtz_x/flame_matrices.py:141-156,199-200).
//
// Scatter is colour-ordered: cells of one colour share no vertex (hence no dof), one
// launch per colour, plain stores -- no atomics, bitwise run-to-run reproducible.
#include <mutex>
#include <vector>

#include "hx_common.cuh"

namespace hx {

// ---------------------------------------------------------------------------------
// reference-element tables (unit-measure scaling), built on the host at first use
// from exact monomial integrals  int_K L^alpha / |K| = d! prod(alpha_i!) / (|alpha|+d)!
// ---------------------------------------------------------------------------------
struct RefTables {
    double M[10][10];      // int phi_a phi_b            (tet)
    double F0[10];         // int phi_a
    double F01[10][4];     // int phi_a L_m
    double F1[10][4][4];   // int phi_a L_m L_n
    double T0[6][6];       // int_F phi_i phi_j          (triangle)
    double T1[6][6][3];    // int_F phi_i phi_j L_m
};
__constant__ RefTables c_ref[2];   // [degree-1]

// 14-point degree-5 rule on the tetrahedron (barycentric), weights sum to 1
__constant__ double c_q14_L[14][4];
__constant__ double c_q14_w[14];

namespace {

struct Term { double c; int e[4]; };
typedef std::vector<Term> Poly;

Poly pmul(const Poly& a, const Poly& b) {
    Poly r;
    for (const Term& x : a)
        for (const Term& y : b) {
            Term t; t.c = x.c * y.c;
            for (int i = 0; i < 4; ++i) t.e[i] = x.e[i] + y.e[i];
            r.push_back(t);
        }
    return r;
}
double fact(int n) { double r = 1; for (int i = 2; i <= n; ++i) r *= i; return r; }
double pint(const Poly& p, int d) {   // d = simplex dimension, uses the first d+1 variables
    double s = 0;
    for (const Term& t : p) {
        int tot = 0; double num = fact(d);
        for (int i = 0; i <= d; ++i) { tot += t.e[i]; num *= fact(t.e[i]); }
        s += t.c * num / fact(tot + d);
    }
    return s;
}
Poly lin(int m) { Term t; t.c = 1; for (int i = 0; i < 4; ++i) t.e[i] = 0; t.e[m] = 1; return Poly{t}; }
std::vector<Poly> basis(int degree, int nv) {
    static const int tet_e[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
    static const int tri_e[3][2] = {{0, 1}, {0, 2}, {1, 2}};
    std::vector<Poly> b;
    if (degree == 1) { for (int a = 0; a < nv; ++a) b.push_back(lin(a)); return b; }
    for (int a = 0; a < nv; ++a) {
        Poly p = pmul(lin(a), lin(a));
        p[0].c = 2.0;
        Poly l = lin(a); l[0].c = -1.0;
        p.push_back(l[0]);
        b.push_back(p);
    }
    const int ne = (nv == 4) ? 6 : 3;
    for (int e = 0; e < ne; ++e) {
        const int* ed = (nv == 4) ? tet_e[e] : tri_e[e];
        Poly p = pmul(lin(ed[0]), lin(ed[1]));
        p[0].c = 4.0;
        b.push_back(p);
    }
    return b;
}

std::once_flag g_tables_once;
int g_tables_status = HX_OK;

void build_tables() {
    RefTables h[2];
    memset(h, 0, sizeof(h));
    for (int deg = 1; deg <= 2; ++deg) {
        RefTables& T = h[deg - 1];
        std::vector<Poly> bt = basis(deg, 4), bf = basis(deg, 3);
        const int nd = (int)bt.size(), nf = (int)bf.size();
        for (int a = 0; a < nd; ++a) {
            T.F0[a] = pint(bt[a], 3);
            for (int b = 0; b < nd; ++b) T.M[a][b] = pint(pmul(bt[a], bt[b]), 3);
            for (int m = 0; m < 4; ++m) {
                Poly am = pmul(bt[a], lin(m));
                T.F01[a][m] = pint(am, 3);
                for (int n = 0; n < 4; ++n) T.F1[a][m][n] = pint(pmul(am, lin(n)), 3);
            }
        }
        for (int i = 0; i < nf; ++i)
            for (int j = 0; j < nf; ++j) {
                Poly ij = pmul(bf[i], bf[j]);
                T.T0[i][j] = pint(ij, 2);
                for (int m = 0; m < 3; ++m) T.T1[i][j][m] = pint(pmul(ij, lin(m)), 2);
            }
    }
    cudaError_t e = cudaMemcpyToSymbol(c_ref, h, sizeof(h));
    // Walkington / Keast 14-point, degree 5
    const double a1 = 0.31088591926330060980, w1 = 0.11268792571801585080;
    const double a2 = 0.092735250310891226402, w2 = 0.073493043116361949544;
    const double b3 = 0.045503704125649649492, w3 = 0.042546020777081466438;
    double L[14][4], w[14];
    int q = 0;
    for (int s = 0; s < 2; ++s) {
        const double a = s ? a2 : a1, ww = s ? w2 : w1;
        for (int i = 0; i < 4; ++i, ++q) {
            for (int k = 0; k < 4; ++k) L[q][k] = a;
            L[q][i] = 1.0 - 3.0 * a;
            w[q] = ww;
        }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = i + 1; j < 4; ++j, ++q) {
            for (int k = 0; k < 4; ++k) L[q][k] = 0.5 - b3;
            L[q][i] = b3; L[q][j] = b3;
            w[q] = w3;
        }
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_q14_L, L, sizeof(L));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_q14_w, w, sizeof(w));
    if (e != cudaSuccess) g_tables_status = fail(HX_ERR_CUDA, "reference tables upload: %s%s", cudaGetErrorString(e));
}

int ensure_tables() {
    std::call_once(g_tables_once, build_tables);
    if (g_tables_status != HX_OK) fail(HX_ERR_CUDA, "reference tables unavailable (no CUDA device?)%s%s");
    return g_tables_status;
}

}  // namespace

// ---------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------
struct TetGeom { double G[4][3]; double vol; double x0[3]; };

__device__ __forceinline__ TetGeom tet_geometry(const double* __restrict__ x, const int* __restrict__ cell) {
    TetGeom g;
    const int v0 = cell[0], v1 = cell[1], v2 = cell[2], v3 = cell[3];
    const double ax = x[3 * v0], ay = x[3 * v0 + 1], az = x[3 * v0 + 2];
    const double e1x = x[3 * v1] - ax, e1y = x[3 * v1 + 1] - ay, e1z = x[3 * v1 + 2] - az;
    const double e2x = x[3 * v2] - ax, e2y = x[3 * v2 + 1] - ay, e2z = x[3 * v2 + 2] - az;
    const double e3x = x[3 * v3] - ax, e3y = x[3 * v3 + 1] - ay, e3z = x[3 * v3 + 2] - az;
    // cofactors: G_1 = e2 x e3 / det, G_2 = e3 x e1 / det, G_3 = e1 x e2 / det
    const double c1x = e2y * e3z - e2z * e3y, c1y = e2z * e3x - e2x * e3z, c1z = e2x * e3y - e2y * e3x;
    const double c2x = e3y * e1z - e3z * e1y, c2y = e3z * e1x - e3x * e1z, c2z = e3x * e1y - e3y * e1x;
    const double c3x = e1y * e2z - e1z * e2y, c3y = e1z * e2x - e1x * e2z, c3z = e1x * e2y - e1y * e2x;
    const double det = e1x * c1x + e1y * c1y + e1z * c1z;
    const double id = 1.0 / det;
    g.G[1][0] = c1x * id; g.G[1][1] = c1y * id; g.G[1][2] = c1z * id;
    g.G[2][0] = c2x * id; g.G[2][1] = c2y * id; g.G[2][2] = c2z * id;
    g.G[3][0] = c3x * id; g.G[3][1] = c3y * id; g.G[3][2] = c3z * id;
    for (int k = 0; k < 3; ++k) g.G[0][k] = -(g.G[1][k] + g.G[2][k] + g.G[3][k]);
    g.vol = fabs(det) / 6.0;
    g.x0[0] = ax; g.x0[1] = ay; g.x0[2] = az;
    return g;
}

__device__ __forceinline__ void tet_edge(int e, int& p, int& q) {
    // (01,02,03,12,13,23)
    p = (e < 3) ? 0 : ((e < 5) ? 1 : 2);
    q = (e < 3) ? e + 1 : ((e < 5) ? e - 1 : 3);
}

// gradient of P2 basis function a at barycentric point L
__device__ __forceinline__ void p2_grad(const TetGeom& g, const double* L, int a, double* out) {
    if (a < 4) {
        const double s = 4.0 * L[a] - 1.0;
        out[0] = s * g.G[a][0]; out[1] = s * g.G[a][1]; out[2] = s * g.G[a][2];
    } else {
        int p, q; tet_edge(a - 4, p, q);
        const double sp = 4.0 * L[p], sq = 4.0 * L[q];
        out[0] = sp * g.G[q][0] + sq * g.G[p][0];
        out[1] = sp * g.G[q][1] + sq * g.G[p][1];
        out[2] = sp * g.G[q][2] + sq * g.G[p][2];
    }
}

__device__ __forceinline__ int find_col(const int* __restrict__ indices, int lo, int hi, int col) {
    // binary search in the sorted row [lo,hi)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int c = indices[mid];
        if (c < col) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---------------------------------------------------------------------------------
// K1/K2: A and C.  One thread per (cell of this colour, local row a).
// ---------------------------------------------------------------------------------
template <int DEG>
__global__ void __launch_bounds__(256)
assemble_AC_kernel(long long n_col_cells, const int* __restrict__ color_cells, const double* __restrict__ x,
                   const int* __restrict__ cells, const int* __restrict__ cell_dofs, const double* __restrict__ cf,
                   int c_is_dg0, const int* __restrict__ indptr, const int* __restrict__ indices,
                   double* __restrict__ a_vals, double* __restrict__ c_vals) {
    constexpr int ND = (DEG == 1) ? 4 : 10;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_col_cells * ND) return;
    const int cell = color_cells[t / ND];
    const int a = (int)(t % ND);
    const int* cv = cells + 4LL * cell;
    const TetGeom g = tet_geometry(x, cv);
    double cn[4];
    double c2_const = 0.0;
    if (c_is_dg0) { const double c = cf[cell]; c2_const = c * c; }
    else { for (int m = 0; m < 4; ++m) cn[m] = cf[cv[m]]; }

    double Arow[ND];
    if (DEG == 1) {
        double cint;
        if (c_is_dg0) cint = g.vol * c2_const;
        else {
            const double s = cn[0] + cn[1] + cn[2] + cn[3];
            const double s2 = cn[0] * cn[0] + cn[1] * cn[1] + cn[2] * cn[2] + cn[3] * cn[3];
            cint = g.vol * (s * s + s2) / 20.0;
        }
#pragma unroll
        for (int b = 0; b < ND; ++b)
            Arow[b] = -cint * (g.G[a][0] * g.G[b][0] + g.G[a][1] * g.G[b][1] + g.G[a][2] * g.G[b][2]);
    } else {
#pragma unroll
        for (int b = 0; b < ND; ++b) Arow[b] = 0.0;
        for (int q = 0; q < 14; ++q) {
            const double* L = c_q14_L[q];
            double c2 = c2_const;
            if (!c_is_dg0) { const double c = cn[0] * L[0] + cn[1] * L[1] + cn[2] * L[2] + cn[3] * L[3]; c2 = c * c; }
            const double wq = -g.vol * c_q14_w[q] * c2;
            double ga[3]; p2_grad(g, L, a, ga);
#pragma unroll
            for (int b = 0; b < ND; ++b) {
                double gb[3]; p2_grad(g, L, b, gb);
                Arow[b] = fma(wq, ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2], Arow[b]);
            }
        }
    }
    const int* dofs = cell_dofs + (long long)ND * cell;
    const int row = dofs[a];
    const int lo = indptr[row], hi = indptr[row + 1];
#pragma unroll
    for (int b = 0; b < ND; ++b) {
        const int p = find_col(indices, lo, hi, dofs[b]);
        a_vals[p] += Arow[b];
        c_vals[p] += g.vol * c_ref[DEG - 1].M[a][b];
    }
}

// ---------------------------------------------------------------------------------
// K3: boundary mass with impedance coefficient. One thread per (facet, local row i).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double tri_area(const double* __restrict__ x, const int* __restrict__ f) {
    const double ax = x[3 * f[0]], ay = x[3 * f[0] + 1], az = x[3 * f[0] + 2];
    const double ux = x[3 * f[1]] - ax, uy = x[3 * f[1] + 1] - ay, uz = x[3 * f[1] + 2] - az;
    const double vx = x[3 * f[2]] - ax, vy = x[3 * f[2] + 1] - ay, vz = x[3 * f[2] + 2] - az;
    const double cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
    return 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
}

template <int DEG>
__global__ void __launch_bounds__(256)
assemble_B_kernel(long long n_col_facets, const int* __restrict__ color_facets, const double* __restrict__ x,
                  const int* __restrict__ facets, const int* __restrict__ facet_dofs,
                  const int* __restrict__ facet_cell, const double* __restrict__ cf, int c_is_dg0, double2 coef,
                  const int* __restrict__ indptr, const int* __restrict__ indices, double2* __restrict__ b_vals) {
    constexpr int NF = (DEG == 1) ? 3 : 6;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_col_facets * NF) return;
    const int fct = color_facets[t / NF];
    const int i = (int)(t % NF);
    const int* fv = facets + 3LL * fct;
    const double area = tri_area(x, fv);
    const int* dofs = facet_dofs + (long long)NF * fct;
    const int row = dofs[i];
    const int lo = indptr[row], hi = indptr[row + 1];
    double c0 = 0, c1 = 0, c2 = 0;
    if (c_is_dg0) c0 = cf[facet_cell[fct]];
    else { c0 = cf[fv[0]]; c1 = cf[fv[1]]; c2 = cf[fv[2]]; }
#pragma unroll
    for (int j = 0; j < NF; ++j) {
        double v;
        if (c_is_dg0) v = c0 * c_ref[DEG - 1].T0[i][j];
        else v = c0 * c_ref[DEG - 1].T1[i][j][0] + c1 * c_ref[DEG - 1].T1[i][j][1] + c2 * c_ref[DEG - 1].T1[i][j][2];
        v *= area;
        const int p = find_col(indices, lo, hi, dofs[j]);
        double2 old = b_vals[p];
        old.x += coef.x * v; old.y += coef.y * v;
        b_vals[p] = old;
    }
}

template <typename T>
__global__ void dirichlet_kernel(int n, const int* __restrict__ indptr, const int* __restrict__ indices,
                                 const unsigned char* __restrict__ is_bc, T* __restrict__ vals);
template <>
__global__ void dirichlet_kernel<double>(int n, const int* __restrict__ indptr, const int* __restrict__ indices,
                                         const unsigned char* __restrict__ is_bc, double* __restrict__ vals) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const bool rb = is_bc[row];
    for (int k = indptr[row]; k < indptr[row + 1]; ++k) {
        const int col = indices[k];
        if (rb || is_bc[col]) vals[k] = (col == row) ? 1.0 : 0.0;
    }
}
template <>
__global__ void dirichlet_kernel<double2>(int n, const int* __restrict__ indptr, const int* __restrict__ indices,
                                          const unsigned char* __restrict__ is_bc, double2* __restrict__ vals) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const bool rb = is_bc[row];
    for (int k = indptr[row]; k < indptr[row + 1]; ++k) {
        const int col = indices[k];
        if (rb || is_bc[col]) vals[k] = make_double2((col == row) ? 1.0 : 0.0, 0.0);
    }
}

__global__ void cell_volumes_kernel(long long n_cells, const double* __restrict__ x, const int* __restrict__ cells,
                                    double* __restrict__ vol) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    vol[c] = tet_geometry(x, cells + 4 * c).vol;
}

// sum_f area, sum_f area*mean(f): single block, fixed order (boundary tags are small)
__global__ void __launch_bounds__(1024)
facet_integrals_kernel(long long n_facets, const double* __restrict__ x, const int* __restrict__ facets,
                       const double* __restrict__ f, double* __restrict__ out2) {
    double sa = 0.0, sf = 0.0;
    for (long long i = threadIdx.x; i < n_facets; i += blockDim.x) {
        const int* fv = facets + 3 * i;
        const double ar = tri_area(x, fv);
        sa += ar;
        if (f) sf += ar * (f[fv[0]] + f[fv[1]] + f[fv[2]]) / 3.0;
    }
    __shared__ double s1[32], s2[32];
    sa = warp_sum(sa); sf = warp_sum(sf);
    if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = sa; s2[threadIdx.x >> 5] = sf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += s1[w]; b += s2[w]; }
        out2[0] = a; out2[1] = b;
    }
}

// ---------------------------------------------------------------------------------
// K5: flame vectors
// ---------------------------------------------------------------------------------
template <int DEG>
__global__ void __launch_bounds__(256)
flame_left_kernel(long long n_col_cells, const int* __restrict__ color_cells, const double* __restrict__ x,
                  const int* __restrict__ cells, const int* __restrict__ cell_dofs, const double* __restrict__ gm1,
                  double gm1_const, const double* __restrict__ h, int h_is_dg0, double scale,
                  const int* __restrict__ cell_tags, int tag, double* __restrict__ out) {
    constexpr int ND = (DEG == 1) ? 4 : 10;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_col_cells * ND) return;
    const int cell = color_cells[t / ND];
    if (cell_tags && cell_tags[cell] != tag) return;
    const int a = (int)(t % ND);
    const int* cv = cells + 4LL * cell;
    const double vol = tet_geometry(x, cv).vol;
    const RefTables& R = c_ref[DEG - 1];
    double v = 0.0;
    if (gm1) {
        double gn[4];
        for (int m = 0; m < 4; ++m) gn[m] = gm1[cv[m]];
        if (h_is_dg0) {
            for (int m = 0; m < 4; ++m) v += R.F01[a][m] * gn[m];
            v *= h[cell];
        } else {
            for (int m = 0; m < 4; ++m) {
                double s = 0.0;
                for (int n = 0; n < 4; ++n) s += R.F1[a][m][n] * h[cv[n]];
                v += gn[m] * s;
            }
        }
    } else {
        if (h_is_dg0) v = R.F0[a] * h[cell];
        else for (int m = 0; m < 4; ++m) v += R.F01[a][m] * h[cv[m]];
        v *= gm1_const;
    }
    out[cell_dofs[(long long)ND * cell + a]] += scale * vol * v;
}

template <int DEG>
__global__ void __launch_bounds__(256)
flame_right_kernel(long long n_col_cells, const int* __restrict__ color_cells, const double* __restrict__ x,
                   const int* __restrict__ cells, const int* __restrict__ cell_dofs, const double* __restrict__ w,
                   const double* __restrict__ rho, double* __restrict__ out) {
    constexpr int ND = (DEG == 1) ? 4 : 10;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_col_cells * ND) return;
    const int cell = color_cells[t / ND];
    const int a = (int)(t % ND);
    const int* cv = cells + 4LL * cell;
    const TetGeom g = tet_geometry(x, cv);
    double wn[4], rn[4];
    for (int m = 0; m < 4; ++m) { wn[m] = w[cv[m]]; rn[m] = rho[cv[m]]; }
    double v = 0.0;
    if (DEG == 1) {
        // FFCx estimated degree 2: 4-point rule (b,a,a,a) & permutations, weights 1/4
        const double qa = 0.1381966011250105, qb = 0.5854101966249685;
        double s = 0.0;
        for (int q = 0; q < 4; ++q) {
            double wq = 0.0, rq = 0.0;
            for (int m = 0; m < 4; ++m) { const double L = (m == q) ? qb : qa; wq += wn[m] * L; rq += rn[m] * L; }
            s += 0.25 * wq / rq;
        }
        v = g.G[a][2] * s;
    } else {
        for (int q = 0; q < 14; ++q) {
            const double* L = c_q14_L[q];
            double wq = 0.0, rq = 0.0;
            for (int m = 0; m < 4; ++m) { wq += wn[m] * L[m]; rq += rn[m] * L[m]; }
            double ga[3]; p2_grad(g, L, a, ga);
            v = fma(c_q14_w[q] * ga[2], wq / rq, v);
        }
    }
    out[cell_dofs[(long long)ND * cell + a]] += g.vol * v;
}

// ---------------------------------------------------------------------------------
// K6: point location / evaluation
// ---------------------------------------------------------------------------------
__global__ void locate_points_kernel(long long n_cells, const double* __restrict__ x, const int* __restrict__ cells,
                                     int n_points, const double* __restrict__ pts, double tol, int* __restrict__ owner) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const TetGeom g = tet_geometry(x, cells + 4 * c);
    for (int p = 0; p < n_points; ++p) {
        const double dx = pts[3 * p] - g.x0[0], dy = pts[3 * p + 1] - g.x0[1], dz = pts[3 * p + 2] - g.x0[2];
        const double l1 = g.G[1][0] * dx + g.G[1][1] * dy + g.G[1][2] * dz;
        const double l2 = g.G[2][0] * dx + g.G[2][1] * dy + g.G[2][2] * dz;
        const double l3 = g.G[3][0] * dx + g.G[3][1] * dy + g.G[3][2] * dz;
        const double l0 = 1.0 - l1 - l2 - l3;
        if (fmin(fmin(l0, l1), fmin(l2, l3)) >= -tol) atomicMin(owner + p, (int)c);
    }
}

template <int DEG>
__global__ void point_dphidz_kernel(const double* __restrict__ x, const int* __restrict__ cells, int n_points,
                                    const double* __restrict__ pts, const int* __restrict__ owner,
                                    double* __restrict__ out) {
    constexpr int ND = (DEG == 1) ? 4 : 10;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_points) return;
    const int c = owner[p];
    if (c < 0 || c >= 0x7f000000) { for (int a = 0; a < ND; ++a) out[p * ND + a] = 0.0; return; }
    const TetGeom g = tet_geometry(x, cells + 4LL * c);
    double L[4];
    const double dx = pts[3 * p] - g.x0[0], dy = pts[3 * p + 1] - g.x0[1], dz = pts[3 * p + 2] - g.x0[2];
    for (int k = 1; k < 4; ++k) L[k] = g.G[k][0] * dx + g.G[k][1] * dy + g.G[k][2] * dz;
    L[0] = 1.0 - L[1] - L[2] - L[3];
    for (int a = 0; a < ND; ++a) {
        if (DEG == 1) out[p * ND + a] = g.G[a][2];
        else { double ga[3]; p2_grad(g, L, a, ga); out[p * ND + a] = ga[2]; }
    }
}

// ---------------------------------------------------------------------------------
// (f-1) shape-derivative boundary integral  int (V.n) div(conj(p_adj) c^2 grad p) ds
// (helmholtz_x/shape_derivatives.py:12-37).  One thread per facet, Dunavant 7-point
// degree-5 rule (exact for P2 p, P1 c, P1 V), single CTA, fixed-order reduction.
// ---------------------------------------------------------------------------------
template <int DEG>
__global__ void __launch_bounds__(1024)
shape_derivative_kernel(int n_sel, const int* __restrict__ sel, const double* __restrict__ x,
                        const int* __restrict__ cells, const int* __restrict__ cell_dofs,
                        const int* __restrict__ facets, const int* __restrict__ facet_cell,
                        const double* __restrict__ V, const double2* __restrict__ p, const double2* __restrict__ pa,
                        const double* __restrict__ cn, double2* __restrict__ out) {
    constexpr int ND = (DEG == 1) ? 4 : 10;
    const double qa[7][3] = {{1.0 / 3, 1.0 / 3, 1.0 / 3},
                             {0.059715871789770, 0.470142064105115, 0.470142064105115},
                             {0.470142064105115, 0.059715871789770, 0.470142064105115},
                             {0.470142064105115, 0.470142064105115, 0.059715871789770},
                             {0.797426985353087, 0.101286507323456, 0.101286507323456},
                             {0.101286507323456, 0.797426985353087, 0.101286507323456},
                             {0.101286507323456, 0.101286507323456, 0.797426985353087}};
    const double qw[7] = {0.225, 0.132394152788506, 0.132394152788506, 0.132394152788506,
                          0.125939180544827, 0.125939180544827, 0.125939180544827};
    double2 acc = make_double2(0.0, 0.0);
    for (int t = threadIdx.x; t < n_sel; t += blockDim.x) {
        const int f = sel[t];
        const int* fv = facets + 3LL * f;
        const int cell = facet_cell[f];
        const int* cv = cells + 4LL * cell;
        const TetGeom g = tet_geometry(x, cv);
        // outward unit normal
        const double ax = x[3 * fv[0]], ay = x[3 * fv[0] + 1], az = x[3 * fv[0] + 2];
        const double ux = x[3 * fv[1]] - ax, uy = x[3 * fv[1] + 1] - ay, uz = x[3 * fv[1] + 2] - az;
        const double vx = x[3 * fv[2]] - ax, vy = x[3 * fv[2] + 1] - ay, vz = x[3 * fv[2] + 2] - az;
        double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
        const double nn = sqrt(nx * nx + ny * ny + nz * nz);
        const double area = 0.5 * nn;
        nx /= nn; ny /= nn; nz /= nn;
        double cx = 0, cy = 0, cz = 0, fx = 0, fy = 0, fz = 0;
        for (int k = 0; k < 4; ++k) { cx += 0.25 * x[3 * cv[k]]; cy += 0.25 * x[3 * cv[k] + 1]; cz += 0.25 * x[3 * cv[k] + 2]; }
        for (int k = 0; k < 3; ++k) { fx += x[3 * fv[k]] / 3; fy += x[3 * fv[k] + 1] / 3; fz += x[3 * fv[k] + 2] / 3; }
        if (nx * (fx - cx) + ny * (fy - cy) + nz * (fz - cz) < 0) { nx = -nx; ny = -ny; nz = -nz; }
        int loc[3];
        for (int k = 0; k < 3; ++k) { loc[k] = 0; for (int a = 0; a < 4; ++a) if (cv[a] == fv[k]) loc[k] = a; }
        const int* dofs = cell_dofs + (long long)ND * cell;
        double cc[4];
        for (int a = 0; a < 4; ++a) cc[a] = cn[cv[a]];
        double gcx = 0, gcy = 0, gcz = 0;
        for (int a = 0; a < 4; ++a) { gcx += cc[a] * g.G[a][0]; gcy += cc[a] * g.G[a][1]; gcz += cc[a] * g.G[a][2]; }
        double2 lap = make_double2(0.0, 0.0);
        if (DEG == 2) {
            for (int a = 0; a < ND; ++a) {
                double l;
                if (a < 4) l = 4.0 * (g.G[a][0] * g.G[a][0] + g.G[a][1] * g.G[a][1] + g.G[a][2] * g.G[a][2]);
                else { int pp, qq; tet_edge(a - 4, pp, qq); l = 8.0 * (g.G[pp][0] * g.G[qq][0] + g.G[pp][1] * g.G[qq][1] + g.G[pp][2] * g.G[qq][2]); }
                const double2 pv = p[dofs[a]];
                lap.x += l * pv.x; lap.y += l * pv.y;
            }
        }
        for (int q = 0; q < 7; ++q) {
            double L[4] = {0, 0, 0, 0};
            for (int k = 0; k < 3; ++k) L[loc[k]] = qa[q][k];
            double2 pq = make_double2(0, 0), aq = make_double2(0, 0);
            double2 gp[3] = {make_double2(0, 0), make_double2(0, 0), make_double2(0, 0)};
            double2 ga[3] = {make_double2(0, 0), make_double2(0, 0), make_double2(0, 0)};
            for (int a = 0; a < ND; ++a) {
                double phi, gr[3];
                if (DEG == 1) { phi = L[a]; gr[0] = g.G[a][0]; gr[1] = g.G[a][1]; gr[2] = g.G[a][2]; }
                else {
                    if (a < 4) phi = L[a] * (2.0 * L[a] - 1.0);
                    else { int pp, qq; tet_edge(a - 4, pp, qq); phi = 4.0 * L[pp] * L[qq]; }
                    p2_grad(g, L, a, gr);
                }
                const double2 pv = p[dofs[a]];
                const double2 av = make_double2(pa[dofs[a]].x, -pa[dofs[a]].y);      // conj(p_adj)
                pq.x += phi * pv.x; pq.y += phi * pv.y;
                aq.x += phi * av.x; aq.y += phi * av.y;
                for (int k = 0; k < 3; ++k) {
                    gp[k].x += gr[k] * pv.x; gp[k].y += gr[k] * pv.y;
                    ga[k].x += gr[k] * av.x; ga[k].y += gr[k] * av.y;
                }
            }
            double cval = 0;
            for (int a = 0; a < 4; ++a) cval += cc[a] * L[a];
            double2 gagp = make_double2(0, 0), gcgp = make_double2(0, 0);
            for (int k = 0; k < 3; ++k) { gagp = cadd(gagp, cmul(ga[k], gp[k])); }
            gcgp.x = gcx * gp[0].x + gcy * gp[1].x + gcz * gp[2].x;
            gcgp.y = gcx * gp[0].y + gcy * gp[1].y + gcz * gp[2].y;
            double2 div = cscale(cval * cval, gagp);
            div = cadd(div, cscale(2.0 * cval, cmul(aq, gcgp)));
            div = cadd(div, cscale(cval * cval, cmul(aq, lap)));
            double vn = 0;
            for (int a = 0; a < 4; ++a)
                vn += L[a] * (V[3 * cv[a]] * nx + V[3 * cv[a] + 1] * ny + V[3 * cv[a] + 2] * nz);
            acc = cadd(acc, cscale(qw[q] * area * vn, div));
        }
    }
    __shared__ double2 sm[32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double2 s = make_double2(0.0, 0.0);
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s = cadd(s, sm[w]);
        out[0] = s;
    }
}

__global__ void threshold_kernel(long long n, double* __restrict__ v, double tol) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && fabs(v[i]) < tol) v[i] = 0.0;
}

// ---------------------------------------------------------------------------------
// K4: sparsity pattern from the cell dofmap
// ---------------------------------------------------------------------------------
__global__ void dof_cell_count_kernel(long long total, const int* __restrict__ cell_dofs, int* __restrict__ count) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < total) atomicAdd(count + cell_dofs[t], 1);
}

__global__ void dof_cell_fill_kernel(long long total, int nd, const int* __restrict__ cell_dofs,
                                     const int* __restrict__ adj_ptr, int* __restrict__ cursor,
                                     int* __restrict__ adj_cells) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int dof = cell_dofs[t];
    const int slot = atomicAdd(cursor + dof, 1);
    adj_cells[adj_ptr[dof] + slot] = (int)(t / nd);
}

// ---- deterministic colouring on the device (Jones-Plassmann with fixed hashed priorities) ----------
// Cells (or facets) conflict when they share a vertex.  Per round every uncoloured cell whose
// priority beats all its uncoloured neighbours takes the smallest colour none of its neighbours
// holds.  Decisions read only the state at the START of the round (ping-pong arrays), so the
// colouring depends on the mesh alone -- not on scheduling -- and the colour-ordered scatter of
// the assembly kernels stays bitwise reproducible.
__device__ __forceinline__ unsigned int color_priority(unsigned int c) {
    unsigned int h = c * 0x9E3779B1u;
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

__global__ void __launch_bounds__(256)
color_round_kernel(long long n_cells, int nv, const int* __restrict__ cells, const int* __restrict__ adj_ptr,
                   const int* __restrict__ adj_cells, const int* __restrict__ cin, int* __restrict__ cout,
                   unsigned long long* __restrict__ remaining) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    int mine = cin[c];
    if (mine == -1) {
        const unsigned int pc = color_priority((unsigned int)c);
        unsigned long long forb0 = 0ull, forb1 = 0ull;
        bool is_max = true;
        for (int k = 0; k < nv && is_max; ++k) {
            const int v = cells[c * nv + k];
            const int s = adj_ptr[v], e = adj_ptr[v + 1];
            for (int t = s; t < e; ++t) {
                const int nb = adj_cells[t];
                if (nb == (int)c) continue;
                const int col = cin[nb];
                if (col < 0) {
                    const unsigned int pn = color_priority((unsigned int)nb);
                    if (pn > pc || (pn == pc && nb > (int)c)) { is_max = false; break; }
                } else if (col < 64) forb0 |= 1ull << col;
                else forb1 |= 1ull << (col - 64);
            }
        }
        if (is_max) {
            if (~forb0) mine = __ffsll((long long)~forb0) - 1;
            else if (~forb1) mine = 64 + __ffsll((long long)~forb1) - 1;
            else mine = -2;                      // more than 128 colours: reported by the host wrapper
        } else {
            atomicAdd(remaining, 1ull);
        }
    }
    cout[c] = mine;
}

// one warp per row: gather the dofs of all incident cells, rank-sort + unique.
constexpr int kPatWarps = 4;
constexpr int kPatCap = 2048;
__global__ void __launch_bounds__(kPatWarps * 32)
pattern_rows_kernel(int n_dofs, int nd, const int* __restrict__ cell_dofs, const int* __restrict__ adj_ptr,
                    const int* __restrict__ adj_cells, int* __restrict__ row_nnz, const int* __restrict__ indptr,
                    int* __restrict__ indices, int write_cols) {
    __shared__ int buf[kPatWarps][kPatCap];
    __shared__ unsigned char flags[kPatWarps][kPatCap];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * kPatWarps + warp;
    if (row >= n_dofs) return;
    int* b = buf[warp];
    const int s = adj_ptr[row], e = adj_ptr[row + 1];
    const int m = (e - s) * nd;
    if (m > kPatCap) { if (lane == 0 && !write_cols) row_nnz[row] = -1; return; }
    for (int t = lane; t < m; t += 32) b[t] = cell_dofs[(long long)adj_cells[s + t / nd] * nd + t % nd];
    __syncwarp();
    // mark duplicates (keep the first occurrence), then rank the survivors
    unsigned char* fl = flags[warp];
    for (int t = lane; t < m; t += 32) {
        const int v = b[t];
        unsigned char first = 1;
        for (int u = 0; u < t; ++u) if (b[u] == v) { first = 0; break; }
        fl[t] = first;
    }
    __syncwarp();
    for (int t = lane; t < m; t += 32) if (!fl[t]) b[t] = 0x7fffffff;
    __syncwarp();
    int cnt = 0;
    const int base = write_cols ? indptr[row] : 0;
    for (int t = lane; t < m; t += 32) {
        const int v = b[t];
        if (v == 0x7fffffff) continue;
        if (!write_cols) { ++cnt; continue; }
        int rank = 0;
        for (int u = 0; u < m; ++u) rank += (b[u] < v);
        indices[base + rank] = v;
    }
    if (!write_cols) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) row_nnz[row] = cnt;
    }
}

}  // namespace hx

using namespace hx;

extern "C" int hx_dof_cell_count(int64_t n_cells, int nd, const int32_t* cell_dofs, int n_dofs, int32_t* count,
                                 hx_stream_t stream) {
    (void)n_dofs;
    const long long total = n_cells * nd;
    if (total <= 0) return HX_OK;
    dof_cell_count_kernel<<<(unsigned)ceil_div<long long>(total, 256), 256, 0, (cudaStream_t)stream>>>(total, cell_dofs, count);
    return check_launch("dof_cell_count_kernel");
}

extern "C" int hx_dof_cell_fill(int64_t n_cells, int nd, const int32_t* cell_dofs, int n_dofs, const int32_t* adj_ptr,
                                int32_t* cursor, int32_t* adj_cells, hx_stream_t stream) {
    (void)n_dofs;
    const long long total = n_cells * nd;
    if (total <= 0) return HX_OK;
    dof_cell_fill_kernel<<<(unsigned)ceil_div<long long>(total, 256), 256, 0, (cudaStream_t)stream>>>(total, nd, cell_dofs, adj_ptr,
                                                                                                  cursor, adj_cells);
    return check_launch("dof_cell_fill_kernel");
}

extern "C" int hx_pattern_rows(int n_dofs, int nd, const int32_t* cell_dofs, const int32_t* adj_ptr, int32_t* adj_cells,
                               int32_t* row_nnz, const int32_t* indptr, int32_t* indices, int write_cols,
                               hx_stream_t stream) {
    if (n_dofs <= 0) return HX_OK;
    pattern_rows_kernel<<<ceil_div(n_dofs, kPatWarps), kPatWarps * 32, 0, (cudaStream_t)stream>>>(
        n_dofs, nd, cell_dofs, adj_ptr, adj_cells, row_nnz, indptr, indices, write_cols);
    return check_launch("pattern_rows_kernel");
}

extern "C" int hx_color_cells(int64_t n_cells, int nv, const int32_t* cells, const int32_t* adj_ptr,
                              const int32_t* adj_cells, int32_t* color_a, int32_t* color_b, int rounds,
                              uint64_t* remaining_dev, hx_stream_t stream) {
    if (n_cells <= 0) return HX_OK;
    if (rounds < 2 || (rounds & 1)) return fail(HX_ERR_ARG, "hx_color_cells: rounds must be even and >= 2%s%s");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)ceil_div<long long>(n_cells, 256);
    int32_t* cin = color_a;
    int32_t* cout = color_b;
    for (int r = 0; r < rounds; ++r) {
        HX_CUDA(cudaMemsetAsync(remaining_dev, 0, sizeof(uint64_t), st));
        color_round_kernel<<<blocks, 256, 0, st>>>(n_cells, nv, cells, adj_ptr, adj_cells, cin, cout,
                                                  (unsigned long long*)remaining_dev);
        int rc = check_launch("color_round_kernel");
        if (rc) return rc;
        int32_t* t = cin; cin = cout; cout = t;
    }
    return HX_OK;            // result in color_a; *remaining_dev = cells still uncoloured
}

extern "C" int hx_color_cells_h(int64_t n_cells, int nv, const int32_t* cells_h, int n_nodes, int32_t* color_h) {
    // greedy first-fit in cell order; per-node bitmask of colours already used around the node
    constexpr int W = 4;   // 256 colours
    std::vector<uint64_t> used((size_t)n_nodes * W, 0);
    int ncol = 0;
    for (int64_t c = 0; c < n_cells; ++c) {
        uint64_t forb[W] = {0, 0, 0, 0};
        const int32_t* cv = cells_h + c * nv;
        for (int k = 0; k < nv; ++k) {
            if (cv[k] < 0 || cv[k] >= n_nodes) return fail(HX_ERR_ARG, "hx_color_cells_h: vertex out of range%s%s");
            for (int w = 0; w < W; ++w) forb[w] |= used[(size_t)cv[k] * W + w];
        }
        int col = -1;
        for (int w = 0; w < W && col < 0; ++w)
            if (~forb[w]) col = w * 64 + __builtin_ctzll(~forb[w]);
        if (col < 0) return fail(HX_ERR_CAPACITY, "hx_color_cells_h: more than 256 colours needed%s%s");
        color_h[c] = col;
        if (col + 1 > ncol) ncol = col + 1;
        for (int k = 0; k < nv; ++k) used[(size_t)cv[k] * W + col / 64] |= (1ULL << (col % 64));
    }
    return ncol;
}

extern "C" int hx_assemble_AC(int degree, int64_t n_cells, const double* x, const int32_t* cells, const int32_t* cell_dofs,
                              const double* c_field, int c_is_dg0, int n_colors, const int64_t* color_ptr_h,
                              const int32_t* color_cells, const int32_t* indptr, const int32_t* indices, double* a_vals,
                              double* c_vals, hx_stream_t stream) {
    (void)n_cells;
    if (degree != 1 && degree != 2) return fail(HX_ERR_ARG, "hx_assemble_AC: degree must be 1 or 2%s%s");
    int rc = ensure_tables();
    if (rc) return rc;
    const int nd = degree == 1 ? 4 : 10;
    for (int c = 0; c < n_colors; ++c) {
        const long long nc = color_ptr_h[c + 1] - color_ptr_h[c];
        if (nc <= 0) continue;
        const unsigned blocks = (unsigned)ceil_div<long long>(nc * nd, 256);
        const int32_t* cc = color_cells + color_ptr_h[c];
        if (degree == 1)
            assemble_AC_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(nc, cc, x, cells, cell_dofs, c_field, c_is_dg0, indptr,
                                                                        indices, a_vals, c_vals);
        else
            assemble_AC_kernel<2><<<blocks, 256, 0, (cudaStream_t)stream>>>(nc, cc, x, cells, cell_dofs, c_field, c_is_dg0, indptr,
                                                                        indices, a_vals, c_vals);
        rc = check_launch("assemble_AC_kernel");
        if (rc) return rc;
    }
    return HX_OK;
}

extern "C" int hx_assemble_B(int degree, int64_t n_facets, const double* x, const int32_t* facets, const int32_t* facet_dofs,
                             const int32_t* facet_cell, const double* c_field, int c_is_dg0, const double* coef_h,
                             int n_colors, const int64_t* color_ptr_h, const int32_t* color_facets, const int32_t* indptr,
                             const int32_t* indices, double* b_vals, hx_stream_t stream) {
    (void)n_facets;
    if (degree != 1 && degree != 2) return fail(HX_ERR_ARG, "hx_assemble_B: degree must be 1 or 2%s%s");
    if (c_is_dg0 && !facet_cell) return fail(HX_ERR_ARG, "hx_assemble_B: facet_cell required for DG0 c%s%s");
    int rc = ensure_tables();
    if (rc) return rc;
    const int nf = degree == 1 ? 3 : 6;
    for (int c = 0; c < n_colors; ++c) {
        const long long nc = color_ptr_h[c + 1] - color_ptr_h[c];
        if (nc <= 0) continue;
        const unsigned blocks = (unsigned)ceil_div<long long>(nc * nf, 256);
        const int32_t* cc = color_facets + color_ptr_h[c];
        if (degree == 1)
            assemble_B_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(nc, cc, x, facets, facet_dofs, facet_cell, c_field,
                                                                       c_is_dg0, h2c(coef_h), indptr, indices, (double2*)b_vals);
        else
            assemble_B_kernel<2><<<blocks, 256, 0, (cudaStream_t)stream>>>(nc, cc, x, facets, facet_dofs, facet_cell, c_field,
                                                                       c_is_dg0, h2c(coef_h), indptr, indices, (double2*)b_vals);
        rc = check_launch("assemble_B_kernel");
        if (rc) return rc;
    }
    return HX_OK;
}

extern "C" int hx_apply_dirichlet(int n, const int32_t* indptr, const int32_t* indices, const uint8_t* is_bc, double* vals,
                                  int is_complex, hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    if (is_complex)
        dirichlet_kernel<double2><<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(n, indptr, indices, is_bc, (double2*)vals);
    else
        dirichlet_kernel<double><<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(n, indptr, indices, is_bc, vals);
    return check_launch("dirichlet_kernel");
}

extern "C" int hx_facet_integrals(int64_t n_facets, const double* x, const int32_t* facets, const double* f_nodal,
                                  double* out2, void* scratch, hx_stream_t stream) {
    (void)scratch;
    facet_integrals_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(n_facets, x, facets, f_nodal, out2);
    return check_launch("facet_integrals_kernel");
}

extern "C" int hx_cell_volumes(int64_t n_cells, const double* x, const int32_t* cells, double* vol, hx_stream_t stream) {
    if (n_cells <= 0) return HX_OK;
    cell_volumes_kernel<<<(unsigned)ceil_div<long long>(n_cells, 256), 256, 0, (cudaStream_t)stream>>>(n_cells, x, cells, vol);
    return check_launch("cell_volumes_kernel");
}

extern "C" int hx_flame_left(int degree, int64_t n_cells, const double* x, const int32_t* cells, const int32_t* cell_dofs,
                             const double* gm1_nodal, double gm1_const, const double* h, int h_is_dg0, double scale,
                             const int32_t* cell_tags, int tag, int n_colors, const int64_t* color_ptr_h,
                             const int32_t* color_cells, double* out, hx_stream_t stream) {
    (void)n_cells;
    if (degree != 1 && degree != 2) return fail(HX_ERR_ARG, "hx_flame_left: degree must be 1 or 2%s%s");
    int rc = ensure_tables();
    if (rc) return rc;
    const int nd = degree == 1 ? 4 : 10;
    for (int c = 0; c < n_colors; ++c) {
        const long long nc = color_ptr_h[c + 1] - color_ptr_h[c];
        if (nc <= 0) continue;
        const unsigned blocks = (unsigned)ceil_div<long long>(nc * nd, 256);
        const int32_t* cc = color_cells + color_ptr_h[c];
        if (degree == 1)
            flame_left_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(nc, cc, x, cells, cell_dofs, gm1_nodal, gm1_const, h,
                                                                       h_is_dg0, scale, cell_tags, tag, out);
        else
            flame_left_kernel<2><<<blocks, 256, 0, (cudaStream_t)stream>>>(nc, cc, x, cells, cell_dofs, gm1_nodal, gm1_const, h,
                                                                       h_is_dg0, scale, cell_tags, tag, out);
        rc = check_launch("flame_left_kernel");
        if (rc) return rc;
    }
    return HX_OK;
}

extern "C" int hx_flame_right(int degree, int64_t n_cells, const double* x, const int32_t* cells, const int32_t* cell_dofs,
                              const double* w_nodal, const double* rho_nodal, int n_colors, const int64_t* color_ptr_h,
                              const int32_t* color_cells, double* out, hx_stream_t stream) {
    (void)n_cells;
    if (degree != 1 && degree != 2) return fail(HX_ERR_ARG, "hx_flame_right: degree must be 1 or 2%s%s");
    int rc = ensure_tables();
    if (rc) return rc;
    const int nd = degree == 1 ? 4 : 10;
    for (int c = 0; c < n_colors; ++c) {
        const long long nc = color_ptr_h[c + 1] - color_ptr_h[c];
        if (nc <= 0) continue;
        const unsigned blocks = (unsigned)ceil_div<long long>(nc * nd, 256);
        const int32_t* cc = color_cells + color_ptr_h[c];
        if (degree == 1)
            flame_right_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(nc, cc, x, cells, cell_dofs, w_nodal, rho_nodal, out);
        else
            flame_right_kernel<2><<<blocks, 256, 0, (cudaStream_t)stream>>>(nc, cc, x, cells, cell_dofs, w_nodal, rho_nodal, out);
        rc = check_launch("flame_right_kernel");
        if (rc) return rc;
    }
    return HX_OK;
}

extern "C" int hx_locate_points(int64_t n_cells, const double* x, const int32_t* cells, int n_points, const double* points,
                                double tol, int32_t* owner, hx_stream_t stream) {
    if (n_cells <= 0 || n_points <= 0) return HX_OK;
    HX_CUDA(cudaMemsetAsync(owner, 0x7f, sizeof(int32_t) * n_points, (cudaStream_t)stream));   // 0x7f7f7f7f: "none"
    locate_points_kernel<<<(unsigned)ceil_div<long long>(n_cells, 256), 256, 0, (cudaStream_t)stream>>>(n_cells, x, cells, n_points,
                                                                                                    points, tol, owner);
    return check_launch("locate_points_kernel");
}

extern "C" int hx_point_dphidz(int degree, const double* x, const int32_t* cells, int n_points, const double* points,
                               const int32_t* owner, double* out, hx_stream_t stream) {
    if (n_points <= 0) return HX_OK;
    if (degree == 1)
        point_dphidz_kernel<1><<<ceil_div(n_points, 64), 64, 0, (cudaStream_t)stream>>>(x, cells, n_points, points, owner, out);
    else if (degree == 2)
        point_dphidz_kernel<2><<<ceil_div(n_points, 64), 64, 0, (cudaStream_t)stream>>>(x, cells, n_points, points, owner, out);
    else
        return fail(HX_ERR_ARG, "hx_point_dphidz: degree must be 1 or 2%s%s");
    return check_launch("point_dphidz_kernel");
}

extern "C" int hx_shape_derivative(int degree, int n_sel, const int32_t* sel, const double* x, const int32_t* cells,
                                   const int32_t* cell_dofs, const int32_t* facets, const int32_t* facet_cell,
                                   const double* V, const double* p, const double* p_adj, const double* c_nodal,
                                   double* out, hx_stream_t stream) {
    if (degree == 1)
        shape_derivative_kernel<1><<<1, 1024, 0, (cudaStream_t)stream>>>(n_sel, sel, x, cells, cell_dofs, facets, facet_cell, V,
                                                                       (const double2*)p, (const double2*)p_adj, c_nodal, (double2*)out);
    else if (degree == 2)
        shape_derivative_kernel<2><<<1, 1024, 0, (cudaStream_t)stream>>>(n_sel, sel, x, cells, cell_dofs, facets, facet_cell, V,
                                                                       (const double2*)p, (const double2*)p_adj, c_nodal, (double2*)out);
    else
        return fail(HX_ERR_ARG, "hx_shape_derivative: degree must be 1 or 2%s%s");
    return check_launch("shape_derivative_kernel");
}

extern "C" int hx_threshold(int64_t n, double* v, double tol, hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    threshold_kernel<<<(unsigned)ceil_div<long long>(n, 256), 256, 0, (cudaStream_t)stream>>>(n, v, tol);
    return check_launch("threshold_kernel");
}
