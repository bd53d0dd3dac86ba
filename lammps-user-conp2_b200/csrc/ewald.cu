// Ewald-mode k-space: k-vector enumeration (host) and the structure-factor,
// b-extraction and Gram-panel kernels.
//
//   ewald_setup_host  KSpaceModuleEwald::conp_setup + make_kvecs_ewald +
//                     make_ug_from_kvecs          km_ewald.cpp:63-132, 277-381
//   axis_tables       fundamentals/harmonics of sincos_a_ele / sincos_b
//                                                km_ewald.cpp:435-457, 685-724
//   ewald_sfac        KSpaceModuleEwald::sincos_b  km_ewald.cpp:668-780
//   ewald_bextract    bbb_from_sincos_b + slabcorr km_ewald.cpp:789-847
//   ewald_panel       operand of the A-matrix Gram (aaa_from_sincos_a :584-666)
//
// Device layout differs from the reference on purpose: instead of the
// flat/expand tables (csk/snk[N][kcount_flat] + kxy_list/kz_list) each atom
// carries three small axis tables E_a[m] = exp(i m k_a r_a); a k-vector's
// phase is the product E_x[kx] E_y[|ky|]^(*) E_z[|kz|]^(*).  The k list is the
// same half-space set as make_kvecs_ewald, ordered (kx, ky) major / kz minor
// so neighbouring threads read neighbouring table entries.
#include "common.cuh"

#include <algorithm>
#include <cmath>

namespace conp {

namespace {
constexpr double MY_PI = 3.14159265358979323846;

// km_ewald.cpp:277-283
double rms(double g_ewald, int km, double prd, long long natoms, double q2) {
  return 2.0 * q2 * g_ewald / prd * std::sqrt(1.0 / (MY_PI * km * natoms)) *
         std::exp(-MY_PI * MY_PI * km * km / (g_ewald * g_ewald * prd * prd));
}
}  // namespace

void ewald_setup_host(EwaldHost &e, double g_ewald, double accuracy, double q2, long long natoms,
                      const double prd[3], double slab_volfactor) {
  const double xprd = prd[0], yprd = prd[1], zprd_slab = prd[2] * slab_volfactor;
  e.volume = xprd * yprd * zprd_slab;
  e.unitk[0] = 2.0 * MY_PI / xprd;
  e.unitk[1] = 2.0 * MY_PI / yprd;
  e.unitk[2] = 2.0 * MY_PI / zprd_slab;
  if (!(accuracy > 0.0) || !(g_ewald > 0.0) || !(q2 >= 0.0) || natoms <= 0)
    CONP_THROW(CONP_ERR_ARG, "conp_set_ewald: need g_ewald > 0, accuracy > 0, natoms > 0");
  const double prds[3] = {xprd, yprd, zprd_slab};
  int kmaxes[3];
  for (int a = 0; a < 3; ++a) {
    int km = 1;
    double err = rms(g_ewald, km, prds[a], natoms, q2);
    while (err > accuracy) {
      km++;
      err = rms(g_ewald, km, prds[a], natoms, q2);
      if (km > 30000) CONP_THROW(CONP_ERR_ARG, "conp_set_ewald: k-space accuracy unreachable (kmax > 30000)");
    }
    kmaxes[a] = km;
  }
  e.kxmax = kmaxes[0]; e.kymax = kmaxes[1]; e.kzmax = kmaxes[2];
  double unitksq[3];
  for (int a = 0; a < 3; ++a) unitksq[a] = e.unitk[a] * e.unitk[a];
  const double gsqxmx = e.unitk[0] * e.unitk[0] * e.kxmax * e.kxmax;
  const double gsqymx = e.unitk[1] * e.unitk[1] * e.kymax * e.kymax;
  const double gsqzmx = e.unitk[2] * e.unitk[2] * e.kzmax * e.kzmax;
  e.gsqmx = std::max(std::max(gsqxmx, gsqymx), gsqzmx) * 1.00001;

  // Half-space enumeration.  sqk is formed left to right exactly as in
  // make_kvecs_ewald (:301, :324, :349) so the `<= gsqmx` test selects the
  // identical set; zero components add exact zeros.
  for (int a = 0; a < 7; ++a) e.dims[a] = 0;
  e.kx.clear(); e.ky.clear(); e.kz.clear(); e.ug.clear();
  const double g_ewald_sq_inv = 1.0 / (g_ewald * g_ewald);
  const double preu = 4.0 * MY_PI / e.volume;
  double ug_tot = 0.0;
  for (int kx = 0; kx <= e.kxmax; ++kx) {
    for (int ky = (kx == 0 ? 0 : -e.kymax); ky <= e.kymax; ++ky) {
      for (int kz = ((kx == 0 && ky == 0) ? 1 : -e.kzmax); kz <= e.kzmax; ++kz) {
        double sqk = 0.0;
        bool first = true;
        const int comp[3] = {kx, ky, kz};
        for (int a = 0; a < 3; ++a) {
          if (comp[a] == 0) continue;
          const double term = (comp[a] * comp[a]) * unitksq[a];
          sqk = first ? term : sqk + term;
          first = false;
        }
        if (!(sqk <= e.gsqmx)) continue;
        const int nz = (kx != 0) + (ky != 0) + (kz != 0);
        // group bookkeeping in the reference's terms (kcount_dims, :298-359)
        if (nz == 1) {
          e.dims[kx ? 0 : (ky ? 1 : 2)]++;
        } else if (nz == 2) {
          const int grp = (kz == 0) ? 3 : (kx == 0 ? 4 : 5);
          // the reference counts one entry per +- pair
          const bool positive = (kz == 0) ? (ky > 0) : (kz > 0);
          if (positive) e.dims[grp]++;
        } else {
          if (ky > 0 && kz > 0) e.dims[6]++;
        }
        e.kx.push_back((short)kx); e.ky.push_back((short)ky); e.kz.push_back((short)kz);
        // make_ug_from_kvecs :375-379 (note: k*k*unitk*unitk, not unitksq)
        double s2 = kx * kx * e.unitk[0] * e.unitk[0];
        s2 += ky * ky * e.unitk[1] * e.unitk[1];
        s2 += kz * kz * e.unitk[2] * e.unitk[2];
        const double u = preu * std::exp(-0.25 * s2 * g_ewald_sq_inv) / s2;
        e.ug.push_back(u);
        ug_tot += 2 * u;
      }
    }
  }
  e.kcount = (int)e.kx.size();
  e.kcount_flat = e.dims[0] + e.dims[1] + e.dims[2] + 2 * e.dims[3];
  e.kcount_expand = e.dims[4] + e.dims[5] + 2 * e.dims[6];
  e.ug_tot = ug_tot;
}

namespace {

// entry t of atom i: axis a, harmonic m
__global__ void __launch_bounds__(256)
axis_tables_kernel(int n, const double *__restrict__ x, const double *__restrict__ y,
                   const double *__restrict__ z, const PosQ *__restrict__ packed, double ukx, double uky,
                   double ukz, int kxmax, int kymax, int kzmax, double2 *__restrict__ tab) {
  const int T = kxmax + kymax + kzmax + 3;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(gid / T);
  if (i >= n) return;
  const int t = (int)(gid - (long long)i * T);
  double r, uk;
  int m;
  if (t <= kxmax) {
    m = t; uk = ukx; r = packed ? packed[i].x : x[i];
  } else if (t <= kxmax + 1 + kymax) {
    m = t - (kxmax + 1); uk = uky; r = packed ? packed[i].y : y[i];
  } else {
    m = t - (kxmax + kymax + 2); uk = ukz; r = packed ? packed[i].z : z[i];
  }
  double s, c;
  sincos(m * (uk * r), &s, &c);
  tab[(size_t)i * T + t] = make_double2(c, s);
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ double2 phase(const double2 *__restrict__ row, int kxmax, int kymax, int kx, int ky,
                                         int kz) {
  const double2 ex = row[kx];
  double2 ey = row[kxmax + 1 + (ky < 0 ? -ky : ky)];
  double2 ez = row[kxmax + kymax + 2 + (kz < 0 ? -kz : kz)];
  if (ky < 0) ey.y = -ey.y;
  if (kz < 0) ez.y = -ez.y;
  return cmul(cmul(ex, ey), ez);
}

constexpr int SF_THREADS = 128;
constexpr int SF_ATOMS = 256;

// S(k) = sum_j q_j exp(i k.r_j): thread per k-vector, block-y per atom chunk
__global__ void __launch_bounds__(SF_THREADS)
sfac_kernel(int m_atoms, const PosQ *__restrict__ atoms, const double2 *__restrict__ tab, int kxmax, int kymax,
            int kzmax, int kcount, const short *__restrict__ kxs, const short *__restrict__ kys,
            const short *__restrict__ kzs, double *__restrict__ sfac) {
  const int k = blockIdx.x * SF_THREADS + threadIdx.x;
  const int T = kxmax + kymax + kzmax + 3;
  const int j0 = blockIdx.y * SF_ATOMS;
  const int j1 = min(j0 + SF_ATOMS, m_atoms);
  if (k >= kcount) return;
  const int kx = kxs[k], ky = kys[k], kz = kzs[k];
  double re = 0.0, im = 0.0;
  for (int j = j0; j < j1; ++j) {
    const double qj = atoms[j].q;
    if (qj == 0.0) continue;  // km_ewald.cpp:686
    const double2 e = phase(tab + (size_t)j * T, kxmax, kymax, kx, ky, kz);
    re = fma(qj, e.x, re);
    im = fma(qj, e.y, im);
  }
  atomicAdd(sfac + 2 * (size_t)k, re);
  atomicAdd(sfac + 2 * (size_t)k + 1, im);
}

constexpr int BX_WARPS = 8;
constexpr int BX_TILE = 256;

// b_i = -sum_k 2 u_k (cos_ik S_re + sin_ik S_im) - z_i * slabcorr ; b = b_k + b_real
__global__ void __launch_bounds__(BX_WARPS * 32)
bextract_kernel(int row_begin, int row_end, const double2 *__restrict__ etab, int kxmax, int kymax, int kzmax,
                int kcount, const short *__restrict__ kxs, const short *__restrict__ kys,
                const short *__restrict__ kzs, const double *__restrict__ ug, const double *__restrict__ sfac,
                const double *__restrict__ ez, const double *__restrict__ qz_sum, double slab_pref,
                const double *__restrict__ b_real, double *__restrict__ b_kspace, double *__restrict__ b) {
  __shared__ short skx[BX_TILE], sky[BX_TILE], skz[BX_TILE];
  __shared__ double2 sw[BX_TILE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = row_begin + blockIdx.x * BX_WARPS + warp;
  const int T = kxmax + kymax + kzmax + 3;
  const bool active = i < row_end;
  const double2 *row = etab + (size_t)(active ? i : row_begin) * T;
  double acc = 0.0;
  for (int k0 = 0; k0 < kcount; k0 += BX_TILE) {
    __syncthreads();
    const int k = k0 + threadIdx.x;
    if (k < kcount) {
      skx[threadIdx.x] = kxs[k]; sky[threadIdx.x] = kys[k]; skz[threadIdx.x] = kzs[k];
      const double u2 = 2.0 * ug[k];
      sw[threadIdx.x] = make_double2(u2 * sfac[2 * (size_t)k], u2 * sfac[2 * (size_t)k + 1]);
    }
    __syncthreads();
    if (active) {
      const int kn = min(BX_TILE, kcount - k0);
      for (int t = lane; t < kn; t += 32) {
        const double2 e = phase(row, kxmax, kymax, skx[t], sky[t], skz[t]);
        acc = fma(e.x, sw[t].x, acc);
        acc = fma(e.y, sw[t].y, acc);
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    double bk = -acc;
    if (slab_pref != 0.0) bk -= ez[i] * (slab_pref * qz_sum[0]);  // slabcorr km_ewald.cpp:839-846
    b_kspace[i] = bk;
    b[i] = bk + b_real[i];
  }
}

// ---------------------------------------------------------------------------
// Tensor-core form of the two O(M K) / O(N K) sums (large systems).  A k-vector's phase factorises,
// exp(i k.r) = [E_x[kx] E_y[ky]] E_z[kz], and the k list is every (kx,ky) pair times a range of kz, so
//     S(kxy, +-m) = sum_j f_j(kxy) E_z,j[m]^(+-1),   f_j(kxy) = q_j E_x,j[kx] E_y,j[ky]
// is a (pairs x atoms) by (atoms x harmonics) matrix product.  With f = fr + i fi, E_z[m] = c + i s the
// four real products P1 = Fr C^T, P2 = Fi S^T, P3 = Fr S^T, P4 = Fi C^T give both signs of kz:
//     S(+m) = (P1 - P2) + i (P3 + P4),   S(-m) = (P1 + P2) + i (P4 - P3).
// All four are blocks of ONE product [Fr | Fi]^T [C | S] of k-major operands (one row per point charge),
// which is exactly the contraction the A-matrix Gram kernel does (gram.cu, FP64 DMMA): launch_tn_gemm
// with the charge index as the contraction dimension, split over CTAs (and over ranks: every rank sums
// its share of the charges, the partial S(k) are then added -- sfac_reduce, km_ewald.cpp:782-786).
// The dense product also covers the (kxy, m) combinations outside the cut-off sphere (about half);
// they are never read.  b extraction is the transposed problem with W_k = 2 u_k S_k:
//     T_i(kxy) = sum_m [W(kxy,+m) conj(E_z,i[m]) + W(kxy,-m) E_z,i[m]],  b_i = -sum_kxy Re(conj(e_xy,i) T_i)
// again one product, over the (harmonic, cos|sin) index: [Tr | Ti] = [Cz_e ; Sz_e]^T [A1 A3 ; A2 A4].
// ---------------------------------------------------------------------------
// operands of one chunk of point charges, one block per charge (row): A row = [Fr | Fi], B row = [Cz | Sz];
// rows jn .. (padding of the contraction dimension to 16) are zero
__global__ void __launch_bounds__(128)
eg_fill_atoms_kernel(int j0, int jn, const PosQ *__restrict__ atoms, const double2 *__restrict__ tab, int T, int kxmax,
                     int kymax, int nkxy, const short *__restrict__ xk, const short *__restrict__ yk, int nkz1,
                     double *__restrict__ FA, size_t wa, double *__restrict__ ZB, size_t wb) {
  const int jl = blockIdx.x;
  double *fa = FA + (size_t)jl * wa, *zb = ZB + (size_t)jl * wb;
  if (jl >= jn) {
    for (int t = threadIdx.x; t < 2 * nkxy; t += blockDim.x) fa[t] = 0.0;
    for (int m = threadIdx.x; m < 2 * nkz1; m += blockDim.x) zb[m] = 0.0;
    return;
  }
  const int j = j0 + jl;
  const double2 *row = tab + (size_t)j * T;
  const double q = atoms[j].q;  // q == 0 (km_ewald.cpp:686) gives exact zeros
  for (int t = threadIdx.x; t < nkxy; t += blockDim.x) {
    const int ky = yk[t];
    const double2 ex = row[xk[t]];
    double2 ey = row[kxmax + 1 + (ky < 0 ? -ky : ky)];
    if (ky < 0) ey.y = -ey.y;
    fa[t] = q * (ex.x * ey.x - ex.y * ey.y);
    fa[nkxy + t] = q * (ex.x * ey.y + ex.y * ey.x);
  }
  for (int m = threadIdx.x; m < nkz1; m += blockDim.x) {
    const double2 ez = row[kxmax + kymax + 2 + m];
    zb[m] = ez.x;
    zb[nkz1 + m] = ez.y;
  }
}

// static electrode-side operands for the rows [row_begin, row_end): unit-charge e_xy (row-major, read by the
// final reduction) and the E_z tables k-major: ZE[m][i] = cos, ZE[nkz1 + m][i] = sin (rows up to kz16 zero)
__global__ void __launch_bounds__(128)
eg_fill_electrodes_kernel(int row_begin, int nrows, const double2 *__restrict__ etab, int T, int kxmax, int kymax,
                          int nkxy, const short *__restrict__ xk, const short *__restrict__ yk, int nkz1,
                          double *__restrict__ FX, double *__restrict__ ZE, size_t wr) {
  const int il = blockIdx.x;
  if (il >= nrows) return;
  const double2 *row = etab + (size_t)(row_begin + il) * T;
  for (int t = threadIdx.x; t < nkxy; t += blockDim.x) {
    const int ky = yk[t];
    const double2 ex = row[xk[t]];
    double2 ey = row[kxmax + 1 + (ky < 0 ? -ky : ky)];
    if (ky < 0) ey.y = -ey.y;
    FX[(size_t)il * 2 * nkxy + t] = ex.x * ey.x - ex.y * ey.y;
    FX[(size_t)il * 2 * nkxy + nkxy + t] = ex.x * ey.y + ex.y * ey.x;
  }
  for (int m = threadIdx.x; m < nkz1; m += blockDim.x) {
    const double2 ez = row[kxmax + kymax + 2 + m];
    ZE[(size_t)m * wr + il] = ez.x;
    ZE[(size_t)(nkz1 + m) * wr + il] = ez.y;
  }
}

// S(k) of the listed k-vectors out of the product's four blocks; the ksplit partial products are added in
// slice order (deterministic)
__global__ void __launch_bounds__(256)
eg_sfac_gather_kernel(int kcount, const int *__restrict__ kxyof, const short *__restrict__ kzs, int nkxy, int nkz1,
                      size_t wb, int ksplit, size_t slice, const double *__restrict__ P, double *__restrict__ sfac) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= kcount) return;
  const int kz = kzs[k], m = kz < 0 ? -kz : kz, a = kxyof[k];
  double p1 = 0.0, p2 = 0.0, p3 = 0.0, p4 = 0.0;
  for (int z = 0; z < ksplit; ++z) {
    const double *Pz = P + (size_t)z * slice;
    p1 += Pz[(size_t)a * wb + m];
    p3 += Pz[(size_t)a * wb + nkz1 + m];
    p4 += Pz[(size_t)(nkxy + a) * wb + m];
    p2 += Pz[(size_t)(nkxy + a) * wb + nkz1 + m];
  }
  sfac[2 * (size_t)k] = kz >= 0 ? p1 - p2 : p1 + p2;
  sfac[2 * (size_t)k + 1] = kz >= 0 ? p3 + p4 : p4 - p3;
}

// W_k = 2 u_k S_k scattered into the four combinations the transposed product needs (A zeroed before),
// k-major over (harmonic, cos|sin): rows m pair with cos E_z, rows nkz1 + m with sin E_z; columns a give
// Tr, columns nkxy + a give Ti:  A1 = Wr+ + Wr-, A2 = Wi+ - Wi-, A3 = Wi+ + Wi-, A4 = Wr- - Wr+.  An entry
// receives at most two addends (kz = +-m), so the atomic sum does not depend on their order.
__global__ void __launch_bounds__(256)
eg_scatter_w_kernel(int kcount, const int *__restrict__ kxyof, const short *__restrict__ kzs, int nkxy, int nkz1,
                    size_t wa, const double *__restrict__ ug, const double *__restrict__ sfac, double *__restrict__ A) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= kcount) return;
  const int kz = kzs[k], m = kz < 0 ? -kz : kz, a = kxyof[k];
  const double u2 = 2.0 * ug[k];
  const double wr = u2 * sfac[2 * (size_t)k], wi = u2 * sfac[2 * (size_t)k + 1];
  atomicAdd(A + (size_t)m * wa + a, wr);
  atomicAdd(A + (size_t)(nkz1 + m) * wa + a, kz >= 0 ? wi : -wi);
  atomicAdd(A + (size_t)m * wa + nkxy + a, wi);
  atomicAdd(A + (size_t)(nkz1 + m) * wa + nkxy + a, kz >= 0 ? -wr : wr);
}

// b_i = -sum_kxy (e_xy,r Tr + e_xy,i Ti) - z_i * slabcorr ; b = b_k + b_real.  One warp per row.
__global__ void __launch_bounds__(256)
eg_b_reduce_kernel(int row_begin, int nrows, int nkxy, const double *__restrict__ FX, const double *__restrict__ Tm,
                   size_t wa, const double *__restrict__ ez, const double *__restrict__ qz_sum, double slab_pref,
                   const double *__restrict__ b_real, double *__restrict__ b_kspace, double *__restrict__ b) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int il = blockIdx.x * 8 + warp;
  if (il >= nrows) return;
  const double *fx = FX + (size_t)il * 2 * nkxy, *tr = Tm + (size_t)il * wa;
  double acc = 0.0;
  for (int t = lane; t < nkxy; t += 32) {
    acc = fma(fx[t], tr[t], acc);
    acc = fma(fx[nkxy + t], tr[nkxy + t], acc);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) {
    const int i = row_begin + il;
    double bk = -acc;
    if (slab_pref != 0.0) bk -= ez[i] * (slab_pref * qz_sum[0]);  // slabcorr km_ewald.cpp:839-846
    b_kspace[i] = bk;
    b[i] = bk + b_real[i];
  }
}

// Gram operand, k-major: Pt[kk][i] = sqrt(2 u_k) cos(k.r_i), Pt[kc+kk][i] = sqrt(2 u_k) sin(k.r_i)
// for k = k0+kk.  One thread per (atom i, segment): a segment is a run of
// consecutive k with equal (kx, ky) and consecutive kz, walked by rotating
// with the z fundamental (the reference builds its harmonics the same way,
// km_ewald.cpp:448-455), restarted from the exact table entry per segment.
__global__ void __launch_bounds__(256)
panel_kernel(int n, const double2 *__restrict__ etab, int kxmax, int kymax, int kzmax, int k0, int kc,
             const int2 *__restrict__ segs, const short *__restrict__ kxs, const short *__restrict__ kys,
             const short *__restrict__ kzs, const double *__restrict__ ug, double *__restrict__ panel, size_t ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int2 sg = segs[blockIdx.y];
  const int T = kxmax + kymax + kzmax + 3;
  const double2 *row = etab + (size_t)i * T;
  double2 e = phase(row, kxmax, kymax, kxs[sg.x], kys[sg.x], kzs[sg.x]);
  const double2 e1 = row[kxmax + kymax + 2 + 1];
  for (int k = sg.x; k < sg.y; ++k) {
    const double s = sqrt(2.0 * ug[k]);
    panel[(size_t)(k - k0) * ld + i] = s * e.x;
    panel[(size_t)(kc + k - k0) * ld + i] = s * e.y;
    e = cmul(e, e1);
  }
}

}  // namespace

int launch_axis_tables(cudaStream_t s, int n, const double *x, const double *y, const double *z,
                       const PosQ *packed, const double unitk[3], int kxmax, int kymax, int kzmax,
                       double2 *tab) {
  if (n <= 0) return 0;
  const long long T = kxmax + kymax + kzmax + 3;
  const long long threads = (long long)n * T;
  axis_tables_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(n, x, y, z, packed, unitk[0], unitk[1],
                                                                      unitk[2], kxmax, kymax, kzmax, tab);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_ewald_sfac(cudaStream_t s, int m, const PosQ *atoms, const double2 *tab, int kxmax, int kymax,
                      int kzmax, int kcount, const short *kx, const short *ky, const short *kz, double *sfac) {
  CUDA_CHECK(cudaMemsetAsync(sfac, 0, sizeof(double) * 2 * (size_t)kcount, s));  // km_ewald.cpp:160-161
  if (m <= 0 || kcount <= 0) return 0;
  dim3 grid((kcount + SF_THREADS - 1) / SF_THREADS, (m + SF_ATOMS - 1) / SF_ATOMS);
  sfac_kernel<<<grid, SF_THREADS, 0, s>>>(m, atoms, tab, kxmax, kymax, kzmax, kcount, kx, ky, kz, sfac);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_ewald_bextract(cudaStream_t s, int row_begin, int row_end, const double2 *etab, int kxmax, int kymax,
                          int kzmax, int kcount, const short *kx, const short *ky, const short *kz,
                          const double *ug, const double *sfac, const double *ez, const double *qz_sum,
                          double slab_pref, const double *b_real, double *b_kspace, double *b) {
  const int n = row_end - row_begin;
  if (n <= 0) return 0;
  bextract_kernel<<<(n + BX_WARPS - 1) / BX_WARPS, BX_WARPS * 32, 0, s>>>(
      row_begin, row_end, etab, kxmax, kymax, kzmax, kcount, kx, ky, kz, ug, sfac, ez, qz_sum, slab_pref, b_real,
      b_kspace, b);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_ewald_panel(cudaStream_t s, int n, const double2 *etab, int kxmax, int kymax, int kzmax, int k0,
                       int kc, int nseg, const int2 *segs, const short *kx, const short *ky, const short *kz,
                       const double *ug, double *panel, size_t ld) {
  if (n <= 0 || kc <= 0 || nseg <= 0) return 0;
  dim3 grid((n + 255) / 256, nseg);
  panel_kernel<<<grid, 256, 0, s>>>(n, etab, kxmax, kymax, kzmax, k0, kc, segs, kx, ky, kz, ug, panel, ld);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

// ---- tensor-core form: host side -------------------------------------------------------------------
namespace {
size_t up_to(size_t v, size_t m) { return (v + m - 1) / m * m; }
}  // namespace

void ewald_gemm_plan(EwaldGemm &g, const EwaldHost &e, int m_total, int nrows, int num_sms, cudaStream_t s) {
  // distinct (kx, ky) pairs in list order (the list is (kx, ky) major, kz minor)
  std::vector<short> xk, yk;
  std::vector<int> kxyof(e.kcount);
  for (int k = 0; k < e.kcount; ++k) {
    if (xk.empty() || xk.back() != e.kx[k] || yk.back() != e.ky[k]) {
      xk.push_back(e.kx[k]);
      yk.push_back(e.ky[k]);
    }
    kxyof[k] = (int)xk.size() - 1;
  }
  g.nkxy = (int)xk.size();
  g.nkz1 = e.kzmax + 1;
  g.np = (size_t)g.nkxy * g.nkz1;
  if (xk.empty()) { xk.push_back(0); yk.push_back(0); }
  if (kxyof.empty()) kxyof.push_back(0);
  g.d_xk.upload(xk, s);
  g.d_yk.upload(yk, s);
  g.d_kxyof.upload(kxyof, s);
  g.wa = up_to(2 * (size_t)std::max(g.nkxy, 1), 128);
  g.wb = up_to(2 * (size_t)g.nkz1, 128);
  g.kz16 = (int)up_to(2 * (size_t)g.nkz1, 16);
  const size_t nr = (size_t)std::max(nrows, 1);
  g.wr = up_to(nr, 128);
  // chunk of point charges whose operands fit ~1 GB; multiple of 16 (contraction tile)
  const size_t per_atom = sizeof(double) * (g.wa + g.wb);
  size_t chunk = ((size_t)1 << 30) / per_atom;
  chunk = std::max<size_t>(1024, chunk / 256 * 256);
  g.chunk = (int)std::min<size_t>(chunk, up_to((size_t)std::max(m_total, 1), 16));
  // split the charge dimension over CTAs so that the grid is one full wave (the tile kernel runs one CTA
  // per SM), with at least 8 contraction tiles per slice
  const size_t tiles = (g.wa / 128) * (g.wb / 128);
  const int want = (int)((size_t)num_sms / tiles);
  g.ksplit = std::max(1, std::min(std::min(want, 64), std::max(g.chunk / 128, 1)));
  g.pslice = 2 * (size_t)std::max(g.nkxy, 1) * g.wb;
  g.d_fa.zero((size_t)g.chunk * g.wa, s);
  g.d_zb.zero((size_t)g.chunk * g.wb, s);
  g.d_p.zero((size_t)g.ksplit * g.pslice, s);
  g.d_a.zero((size_t)g.kz16 * g.wa, s);
  g.d_ze.zero((size_t)g.kz16 * g.wr, s);
  g.d_t.zero(nr * g.wa, s);
  g.d_fx.zero(nr * 2 * (size_t)std::max(g.nkxy, 1), s);
}

int ewald_gemm_electrodes(cudaStream_t s, EwaldGemm &g, const EwaldHost &e, int row_begin, int row_end,
                          const double2 *etab) {
  const int n = row_end - row_begin;
  if (n <= 0 || g.nkxy <= 0) return 0;
  const int T = e.kxmax + e.kymax + e.kzmax + 3;
  eg_fill_electrodes_kernel<<<n, 128, 0, s>>>(row_begin, n, etab, T, e.kxmax, e.kymax, g.nkxy, g.d_xk.p, g.d_yk.p,
                                              g.nkz1, g.d_fx.p, g.d_ze.p, g.wr);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int ewald_gemm_sfac(cudaStream_t s, EwaldGemm &g, const EwaldHost &e, int j_begin, int j_end, const PosQ *atoms,
                    const double2 *tab, const short *kz, double *sfac) {
  if (e.kcount <= 0) return 0;
  const int m = j_end - j_begin;
  if (m <= 0) {
    CUDA_CHECK(cudaMemsetAsync(sfac, 0, sizeof(double) * 2 * (size_t)e.kcount, s));  // km_ewald.cpp:160-161
    return 0;
  }
  const int T = e.kxmax + e.kymax + e.kzmax + 3;
  int launched = 0;
  for (int j0 = 0; j0 < m; j0 += g.chunk) {
    const int jn = std::min(g.chunk, m - j0);
    const int jn16 = (int)up_to(jn, 16);
    eg_fill_atoms_kernel<<<jn16, 128, 0, s>>>(j_begin + j0, jn, atoms, tab, T, e.kxmax, e.kymax, g.nkxy, g.d_xk.p,
                                              g.d_yk.p, g.nkz1, g.d_fa.p, g.wa, g.d_zb.p, g.wb);
    CUDA_CHECK(cudaGetLastError());
    // P[z] (2 nkxy x 2 nkz1) (+)= [Fr | Fi]^T [Cz | Sz] over the charges of slice z
    launched += 1 + launch_tn_gemm(s, 2 * g.nkxy, 2 * g.nkz1, jn16, g.d_fa.p, g.wa, g.d_zb.p, g.wb, g.d_p.p, g.wb,
                                   g.ksplit, g.pslice, j0 > 0);
  }
  eg_sfac_gather_kernel<<<(e.kcount + 255) / 256, 256, 0, s>>>(e.kcount, g.d_kxyof.p, kz, g.nkxy, g.nkz1, g.wb,
                                                               g.ksplit, g.pslice, g.d_p.p, sfac);
  CUDA_CHECK(cudaGetLastError());
  return launched + 1;
}

int ewald_gemm_bextract(cudaStream_t s, EwaldGemm &g, const EwaldHost &e, int row_begin, int row_end,
                        const short *kz, const double *ug, const double *sfac, const double *ez,
                        const double *qz_sum, double slab_pref, const double *b_real, double *b_kspace, double *b) {
  const int n = row_end - row_begin;
  if (n <= 0) return 0;
  CUDA_CHECK(cudaMemsetAsync(g.d_a.p, 0, sizeof(double) * (size_t)g.kz16 * g.wa, s));
  int launched = 0;
  if (e.kcount > 0) {
    eg_scatter_w_kernel<<<(e.kcount + 255) / 256, 256, 0, s>>>(e.kcount, g.d_kxyof.p, kz, g.nkxy, g.nkz1, g.wa, ug,
                                                               sfac, g.d_a.p);
    CUDA_CHECK(cudaGetLastError());
    ++launched;
  }
  // [Tr | Ti] (n x 2 nkxy) = [Cz_e ; Sz_e]^T (n x kz16) . A (kz16 x 2 nkxy)
  launched += launch_tn_gemm(s, n, 2 * g.nkxy, g.kz16, g.d_ze.p, g.wr, g.d_a.p, g.wa, g.d_t.p, g.wa, 1, 0, 0);
  eg_b_reduce_kernel<<<(n + 7) / 8, 256, 0, s>>>(row_begin, n, g.nkxy, g.d_fx.p, g.d_t.p, g.wa, ez, qz_sum, slab_pref,
                                                 b_real, b_kspace, b);
  CUDA_CHECK(cudaGetLastError());
  return launched + 1;
}


}  // namespace conp
