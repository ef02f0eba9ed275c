// Device-side DCMIP 2012 test 1-1 (3D deformational flow) and 1-2 (Hadley-like circulation): prescribed winds
// and initial tracers, so that a run keeps every input in HBM (no host round trip per step).
//
// Restates the point functions of reference src/share/dcmip_123_mod.F90:85-272 (1-1) and :279-409 (1-2) and the
// element fill of src/share/dcmip_wrapper_mod.F90:49-266, including the wrapper's quirks that change numbers:
// z coordinates (zcoords=1, z = H ln(1/eta)), dp = p_i(k+1)-p_i(k), tracers beyond the 4 analytic ones replaced by
// the sin(9 lon) sin(9 lat) checkerboard (:215-243), omega_p not defined by the reference (left 0).
// The flow is separable, so everything that depends only on the level is tabulated once on the host
// (DcmipTables) and the kernels evaluate only the horizontal trigonometry per node.
#pragma once
#include <cmath>

#include "tse_layout.cuh"

namespace tse {

struct DcmipConst {
  static constexpr double pi = 3.141592653589793238462643383279;
  static constexpr double a = 6.376e6, Rd = 287.04, g = 9.80616, p0 = 100000.0, T0 = 300.0;
};

// per-level tables (device pointers)
struct DcmipTables {
  const double* zm;      // [72] mid-level height
  const double* dp_ref;  // [72] p_i(k+1)-p_i(k)
  const double* vm;      // [72] vertical factor of the level-dependent wind component (ud for 1-1, v for 1-2)
  const double* vi;      // [72] vertical factor of eta_dot_dpdn = -g*rho*w at interfaces 1..72
  const double* dp_ic;   // [72] dA*ps0 + dB*ps_v(t=0)  (prim_driver_mod.F90:646-669)
};

struct DcmipHostTables {
  double zm[NLEV], zi[NLEV + 1], pm[NLEV], pint[NLEV + 1], dp_ref[NLEV], vm[NLEV], vi[NLEV], dp_ic[NLEV];
};

// host: fill the tables for test 11 / 12 from the hybrid coefficients
inline void dcmip_fill_tables(int test, const double* hyai, const double* hybi, const double* hyam, const double* hybm, double ps0,
                              DcmipHostTables& t) {
  typedef DcmipConst C;
  const double H = C::Rd * C::T0 / C::g;
  for (int k = 0; k < NLEV; ++k) {
    t.zm[k] = H * std::log(1.0 / (hyam[k] + hybm[k]));  // dcmip_wrapper_mod.F90:64,121
    t.pm[k] = C::p0 * std::exp(-t.zm[k] / H);
  }
  for (int k = 0; k <= NLEV; ++k) {
    t.zi[k] = H * std::log(1.0 / (hyai[k] + hybi[k]));
    t.pint[k] = C::p0 * std::exp(-t.zi[k] / H);
  }
  for (int k = 0; k < NLEV; ++k) {
    t.dp_ref[k] = t.pint[k + 1] - t.pint[k];                                         // :183
    t.dp_ic[k] = (hyai[k + 1] - hyai[k]) * ps0 + (hybi[k + 1] - hybi[k]) * t.pint[NLEV];  // ps_v = p_i(nlevp)
  }
  if (test == 11) {
    const double tau = 12.0 * 86400.0, omega0 = (23000.0 * C::pi) / tau;
    const double ptop = C::p0 * std::exp(-12000.0 / H);
    const double bs = (double)0.2f;  // "bs = 0.2" is a default-real literal (dcmip_123_mod.F90:161)
    for (int k = 0; k < NLEV; ++k) {
      const double plim = std::max(t.pm[k], ptop);
      t.vm[k] = (omega0 * C::a) / (bs * ptop) * (-std::exp((plim - C::p0) / (bs * ptop)) + std::exp((ptop - plim) / (bs * ptop)));
      const double p = t.pint[k], pl = std::max(p, ptop);
      const double s = 1.0 + std::exp((ptop - C::p0) / (bs * ptop)) - std::exp((pl - C::p0) / (bs * ptop)) - std::exp((ptop - pl) / (bs * ptop));
      const double rho = p / (C::Rd * C::T0);
      t.vi[k] = -C::g * rho * (-((C::Rd * C::T0) / (C::g * pl)) * omega0 * s);
    }
  } else {
    const double w0 = 0.15, K = 5.0, ztop = 12000.0;
    const double ptop = C::p0 * std::exp(-ztop / H);
    const double rho0 = C::p0 / (C::Rd * C::T0);
    for (int k = 0; k < NLEV; ++k) {
      const double rho = std::max(t.pm[k], ptop) / (C::Rd * C::T0);
      const double hstar = std::min(t.zm[k] / ztop, 1.0);
      t.vm[k] = -(rho0 / rho) * (C::a * w0 * C::pi) / (K * ztop) * std::cos(C::pi * hstar);
      const double rhoi = std::max(t.pint[k], ptop) / (C::Rd * C::T0);
      const double hsi = std::min(t.zi[k] / ztop, 1.0);
      t.vi[k] = -C::g * rhoi * ((rho0 / rhoi) * (w0 / K) * std::sin(C::pi * hsi));
    }
  }
}

// prim_advance_exp for the prescribed-wind cases (prim_advance_mod.F90:111-149): vn0 = v(n0)*dp with v(n0) the wind the
// previous step evaluated (t_prev), derived%dp = dp_ref, eta_dot_dpdn = -g rho w at the current time, omega_p = 0.
// Block = (group, level chunk), thread = (element in group, node).
__global__ void __launch_bounds__(256) k_dcmip_wind(int test, double t_prev, double t_now, int nelem, const double* __restrict__ lon,
                                                    const double* __restrict__ lat, DcmipTables tb, double* __restrict__ vn0,
                                                    double* __restrict__ dp, double* __restrict__ eta_dot, double* __restrict__ omega_p) {
  typedef DcmipConst C;
  const int g = blockIdx.x / NKC, kc = blockIdx.x % NKC;
  const int el = threadIdx.x >> 4, n = threadIdx.x & 15;
  const int e = g * GE + el;
  if (e >= nelem) return;
  const double lo = lon[(size_t)e * 16 + n], la = lat[(size_t)e * 16 + n];
  double hu0, hu1, hv, he;  // u = hu0 + hu1*vm(k) ; v = hv [*vm(k) for 1-2] ; eta_dot = he*vi(k)
  if (test == 11) {
    const double tau = 12.0 * 86400.0, u0 = (2.0 * C::pi * C::a) / tau, k0 = (10.0 * C::a) / tau;
    const double lonp = lo - 2.0 * C::pi * t_prev / tau;
    const double cl = cos(la), sl = sin(lonp);
    hu0 = k0 * sl * sl * sin(2.0 * la) * cos(C::pi * t_prev / tau) + u0 * cl;
    hu1 = cos(lonp) * (cl * cl) * cos(2.0 * C::pi * t_prev / tau);
    hv = k0 * sin(2.0 * lonp) * cl * cos(C::pi * t_prev / tau);
    const double lonn = lo - 2.0 * C::pi * t_now / tau;
    he = sin(lonn) * cl * cos(2.0 * C::pi * t_now / tau);
  } else {
    const double tau = 1.0 * 86400.0, u0 = 40.0, K = 5.0;
    const double cl = cos(la);
    hu0 = u0 * cl;
    hu1 = 0.0;
    hv = cl * sin(K * la) * cos(C::pi * t_prev / tau);
    he = (-2.0 * sin(K * la) * sin(la) + K * cl * cos(K * la)) * cos(C::pi * t_now / tau);
  }
#pragma unroll
  for (int kk = 0; kk < KC; ++kk) {
    const int k = kc * KC + kk;
    const double d = tb.dp_ref[k];
    const double u = hu0 + hu1 * tb.vm[k];
    const double v = (test == 11) ? hv : hv * tb.vm[k];
    const size_t lp = lplane(e, k) * 16 + n;
    vn0[vplane(e, k, 0) * 16 + n] = u * d;
    vn0[vplane(e, k, 1) * 16 + n] = v * d;
    dp[lp] = d;
    eta_dot[lp] = he * tb.vi[k];
    omega_p[lp] = 0.0;
  }
}

// initial tracers (time 0) for both Qdp time levels: Qdp = q * (dA*ps0 + dB*ps_v)  (prim_driver_mod.F90:646-669)
__global__ void __launch_bounds__(256) k_dcmip_ic(int test, int nelem, int Q, const double* __restrict__ lon, const double* __restrict__ lat,
                                                  DcmipTables tb, double* __restrict__ qa, double* __restrict__ qb) {
  typedef DcmipConst C;
  const int g = blockIdx.x / NKC, kc = blockIdx.x % NKC;
  const int el = threadIdx.x >> 4, n = threadIdx.x & 15;
  const int e = g * GE + el;
  if (e >= nelem) return;
  const double lo = lon[(size_t)e * 16 + n], la = lat[(size_t)e * 16 + n];
  const double checker = (sin(9. * lo) * sin(9. * la) < 0.) ? 0.0 : 1.0;  // dcmip_wrapper_mod.F90:215-243
  double r1 = 0, r2 = 0;
  if (test == 11) {
    const double lambda0 = 5.0 * C::pi / 6.0, lambda1 = 7.0 * C::pi / 6.0;
    // phi0 = phi1 = 0: sin_tmp = sin(lat)*sin(0) = 0, cos_tmp = cos(lat)*cos(0)
    const double st = sin(la) * sin(0.0), ct = cos(la) * cos(0.0);
    r1 = acos(st + ct * cos(lo - lambda0));
    r2 = acos(st + ct * cos(lo - lambda1));
  }
  for (int kk = 0; kk < KC; ++kk) {
    const int k = kc * KC + kk;
    const double height = tb.zm[k];
    double q4[4];
    if (test == 11) {
      const double RR = 0.5, ZZ = 1000.0, z0 = 5000.0;
      const double hz = (height - z0) / ZZ;
      const double d1 = fmin(1.0, (r1 / RR) * (r1 / RR) + hz * hz);
      const double d2 = fmin(1.0, (r2 / RR) * (r2 / RR) + hz * hz);
      const double q1 = 0.5 * (1.0 + cos(C::pi * d1)) + 0.5 * (1.0 + cos(C::pi * d2));
      const double q2 = 0.9 - 0.8 * q1 * q1;
      double q3 = (d1 <= RR || d2 <= RR) ? 1.0 : 0.1;
      if (height > z0 && fabs(la) < 0.125) q3 = 0.1;
      q4[0] = q1; q4[1] = q2; q4[2] = q3; q4[3] = 1.0 - 0.3 * (q1 + q2 + q3);
    } else {
      const double z1 = 2000.0, z2 = 5000.0, z0 = 0.5 * (z1 + z2);
      q4[0] = 0.0;
      q4[1] = (height < z2 && height > z1) ? 0.5 * (1.0 + cos(2.0 * C::pi * (height - z0) / (z2 - z1))) : 0.0;
      q4[2] = q4[3] = 0.0;
    }
    const double d = tb.dp_ic[k];
    for (int q = 0; q < Q; ++q) {
      double qv;
      if (test == 11) qv = q < 4 ? q4[q] : checker;
      else qv = (q == 1) ? q4[1] : checker;  // 1-2: tracer 2 is the Hadley layer, 1 and 3.. are the checkerboard
      const size_t idx = qplane(e, q, k, Q) * 16 + n;
      qa[idx] = qv * d;
      qb[idx] = qv * d;
    }
  }
}

}  // namespace tse
