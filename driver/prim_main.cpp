// prim_main: stand-alone driver of the B200 tracer-advection library, following the reference's main program
// (reference src/prim_main.F90:37-203: namelist from stdin, prim_init1/prim_init2, the prim_run_subcycle loop, prim_finalize)
// for the two test cases the mini-app ships (DCMIP 1-1 / 1-2, prescribed winds).
//
//     ./prim_main < test/dcmip1-1/dcmip1-1.nl          (after the sed of the run scripts: NE, TIME_STEP, qsize, nu_q)
//     ./prim_main - ne=8 tstep=400 < dcmip1-1.nl       (stdin namelist plus key=value overrides)
//     ./prim_main ne=8 tstep=400 nu_q=6e16 qsize=4 test_case=dcmip1-1 ndays=12     (key=value settings only)
//
// What it keeps from the reference:
//   * the namelist groups and keys that the path reads (ctl_nl: ne, qsize, ndays, nmax, tstep, qsplit, rsplit, nu_q,
//     limiter_option, hypervis_order, hypervis_subcycle_q, test_case, statefreq, prescribed_wind; vert_nl: vfile_mid, vfile_int --
//     namelist_mod.F90:159-263,688-692); unknown keys and groups are accepted and ignored, as a Fortran namelist read would
//     ignore groups it does not ask for;
//   * the hybrid-coordinate ascii tables (hybvcoord_mod.F90:36-153);
//   * nEndStep = nmax, or ndays*secpday/tstep when ndays > 0 (prim_driver_mod.F90:599-606), the time loop
//     `do while (tl%nstep < nEndStep) call prim_run_subcycle` (prim_main.F90:142-175), prim_printstate's tracer lines
//     (prim_state_mod.F90:352-385: qv min/max, Q mass) every statefreq steps;
//   * the end-of-run verification the run scripts do with NCL (test/dcmip1-*/dcmip1-*_error_norm_ng.ncl): L1/L2/Linf, q_max, q_min.
// What it leaves out: netCDF history/restart output (prim_movie_mod, restart_io_mod), GPTL files, MPI (one process per GPU:
// launch N copies under an MPI/torchrun launcher and hand the rank/size in through TSE_RANK/TSE_NRANKS + an id file, see
// INTEGRATION.md; this driver runs single-GPU).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "tse.h"
#include "tse_mesh.hpp"

namespace {

[[noreturn]] void abortmp(const std::string& msg) {  // parallel_mod.F90:274-290
  std::fprintf(stderr, "ERROR: %s\n", msg.c_str());
  std::exit(1);
}
#define TSE_CALL(call)                                                         \
  do {                                                                         \
    if ((call) != 0) abortmp(std::string(#call) + ": " + tse_last_error());    \
  } while (0)

// ---- namelist (the subset of Fortran namelist syntax the shipped files use) --------------------------------------------------
std::string lower(std::string s) {
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return std::tolower(c); });
  return s;
}
std::string trim(const std::string& s) {
  const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n,");
  return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}
std::map<std::string, std::string> read_namelist(std::istream& in) {
  std::map<std::string, std::string> kv;  // "group.key" -> value (quotes stripped)
  std::string line, group;
  while (std::getline(in, line)) {
    bool inq = false;
    for (size_t i = 0; i < line.size(); ++i) {  // strip the comment (a '!' outside quotes)
      if (line[i] == '\'' || line[i] == '"') inq = !inq;
      if (line[i] == '!' && !inq) {
        line.resize(i);
        break;
      }
    }
    line = trim(line);
    if (line.empty()) continue;
    if (line[0] == '&') {
      group = lower(trim(line.substr(1)));
      continue;
    }
    if (line[0] == '/') {
      group.clear();
      continue;
    }
    const size_t eq = line.find('=');
    if (eq == std::string::npos || group.empty()) continue;
    std::string key = lower(trim(line.substr(0, eq))), val = trim(line.substr(eq + 1));
    if (val.size() >= 2 && (val.front() == '"' || val.front() == '\'')) val = val.substr(1, val.find_last_of("\"'") - 1);
    kv[group + "." + key] = val;
  }
  return kv;
}
double to_double(std::string v) {  // Fortran reals: 6e16, 0.04D0, .80
  for (char& c : v)
    if (c == 'D' || c == 'd') c = 'e';
  return std::strtod(v.c_str(), nullptr);
}

// hybrid coefficient table: "<n> ! name" followed by n values (hybvcoord_mod.F90:84-153)
std::vector<std::vector<double>> read_vcoord(const std::string& path) {
  std::ifstream f(path);
  if (!f) abortmp("cannot open vertical coordinate file " + path);
  std::vector<std::vector<double>> out;
  std::string line;
  while (std::getline(f, line)) {
    line = trim(line.substr(0, line.find('!')));
    if (line.empty()) continue;
    const int n = std::atoi(line.c_str());
    std::vector<double> a;
    while ((int)a.size() < n && std::getline(f, line)) {
      std::replace(line.begin(), line.end(), ',', ' ');
      std::istringstream ss(line);
      std::string tok;
      while (ss >> tok) a.push_back(to_double(tok));
    }
    if ((int)a.size() != n) abortmp("short table in " + path);
    out.push_back(a);
  }
  return out;
}

}  // namespace

int main(int argc, char** argv) {
  // ---- readnl (namelist_mod.F90:266-1039) ----
  std::map<std::string, std::string> nl;
  bool have_file = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "-") continue;
    if (a.find('=') == std::string::npos) {
      std::ifstream f(a);
      if (!f) abortmp("cannot open namelist " + a);
      nl = read_namelist(f);
      have_file = true;
    }
  }
  // like the reference, the namelist comes from stdin -- when no file and no key=value settings are given, or when asked with "-"
  bool want_stdin = argc == 1;
  for (int i = 1; i < argc; ++i) want_stdin |= std::string(argv[i]) == "-";
  if (!have_file && want_stdin) nl = read_namelist(std::cin);
  for (int i = 1; i < argc; ++i) {  // key=value overrides (what the run scripts do with sed)
    const std::string a = argv[i];
    const size_t eq = a.find('=');
    if (eq == std::string::npos) continue;
    if (a == "-") continue;
    const std::string k = lower(a.substr(0, eq));
    nl[(k == "vfile_mid" || k == "vfile_int" ? "vert_nl." : "ctl_nl.") + k] = a.substr(eq + 1);
  }
  auto geti = [&](const char* k, int dflt) { auto it = nl.find(std::string("ctl_nl.") + k); return it == nl.end() ? dflt : (int)to_double(it->second); };
  auto getd = [&](const char* k, double dflt) { auto it = nl.find(std::string("ctl_nl.") + k); return it == nl.end() ? dflt : to_double(it->second); };
  auto gets = [&](const char* k, const char* dflt) { auto it = nl.find(k); return it == nl.end() ? std::string(dflt) : it->second; };

  tse_config cfg{};
  cfg.ne = geti("ne", 0);
  cfg.qsize = geti("qsize", 4);
  cfg.qsize_d = std::max(cfg.qsize, 1);
  cfg.nlev = 72;
  cfg.np = 4;
  cfg.rsplit = geti("rsplit", 3);
  cfg.qsplit = geti("qsplit", 1);
  cfg.limiter_option = geti("limiter_option", 8);
  cfg.hypervis_order = geti("hypervis_order", 2);
  cfg.hypervis_subcycle_q = geti("hypervis_subcycle_q", 1);
  cfg.vert_remap_q_alg = geti("vert_remap_q_alg", 0);
  cfg.nu_q = getd("nu_q", 0.0);
  cfg.device = -1;
  const double tstep = getd("tstep", 0.0);
  const int ndays = geti("ndays", 0), nmax = geti("nmax", 12), statefreq = std::max(1, geti("statefreq", 99999));
  const std::string test_case = lower(gets("ctl_nl.test_case", "dcmip1-1"));
  const int test = test_case == "dcmip1-1" ? 11 : test_case == "dcmip1-2" ? 12 : 0;
  if (cfg.ne <= 0 || tstep <= 0) abortmp("namelist: ne and tstep must be set (ctl_nl)");
  if (!test) abortmp("test_case must be dcmip1-1 or dcmip1-2 (got " + test_case + ")");
  if (geti("prescribed_wind", 1) != 1) abortmp("this driver runs the prescribed-wind transport cases only (prescribed_wind = 1)");
  const int nEndStep = ndays > 0 ? (int)std::llround(ndays * 86400.0 / tstep) : nmax;  // prim_driver_mod.F90:599-606

  const auto vi = read_vcoord(gets("vert_nl.vfile_int", "vcoord/acme-72i.ascii"));
  const auto vm = read_vcoord(gets("vert_nl.vfile_mid", "vcoord/acme-72m.ascii"));
  if (vi.size() < 2 || vm.size() < 2 || vi[0].size() != 73 || vm[0].size() != 72) abortmp("vertical coordinate tables must hold 72 levels");
  tse_hvcoord hv{vi[0].data(), vi[1].data(), vm[0].data(), vm[1].data(), 100000.0};

  // ---- prim_init1: mesh, metric terms, edge descriptors (prim_driver_mod.F90:86-540) ----
  std::printf(" ne = %d  nelem = %d  qsize = %d  tstep = %g  nu_q = %g  test_case = %s  nEndStep = %d\n", cfg.ne, 6 * cfg.ne * cfg.ne, cfg.qsize,
              tstep, cfg.nu_q, test_case.c_str(), nEndStep);
  tse::Mesh mesh(cfg.ne);
  const std::vector<int> owner(mesh.nelem, 0);
  const tse::LocalView view = tse::make_local_view(mesh, owner, 0, 1);
  cfg.nelemd = view.nelemd;
  tse_geometry geom{mesh.spheremp.data(), mesh.rspheremp.data(), mesh.metdet.data(), mesh.rmetdet.data(), mesh.Dinv.data(), mesh.lat.data(), mesh.lon.data()};
  tse_connectivity conn{view.putmap.data(), view.getmap.data(), view.reverse.data(), view.nbuf, mesh.sfc.data(), 0, nullptr, nullptr, nullptr};
  tse_handle h = nullptr;
  TSE_CALL(tse_init(&cfg, &geom, &conn, &hv, tse::gll().dvv, &h));

  // ---- prim_init2: initial state (prim_driver_mod.F90:546-699) ----
  TSE_CALL(tse_dcmip_init(h, test));
  const int Q = cfg.qsize;
  std::vector<double> mass0(Q), mass(Q), qmn(Q), qmx(Q);
  auto printstate = [&](int nstep, int tl) {  // prim_state_mod.F90:352-385
    TSE_CALL(tse_diag_mass(h, tl, mass.data()));
    TSE_CALL(tse_diag_qminmax(h, tl, qmn.data(), qmx.data()));
    std::printf(" nstep= %d  time= %.6f [day]\n", nstep, nstep * tstep / 86400.0);
    for (int q = 0; q < Q; ++q)
      std::printf("   Q%-2d min/max = %23.15e %23.15e   mass = %23.15e   (mass-mass0)/mass0 = %10.3e\n", q + 1, qmn[q], qmx[q], mass[q],
                  mass0[q] != 0 ? (mass[q] - mass0[q]) / mass0[q] : 0.0);
  };
  TSE_CALL(tse_diag_mass(h, 1, mass0.data()));
  printstate(0, 1);
  // t = 0 mixing ratio of the tracer the norms are taken on (q1 for 1-1, q2 for 1-2), Q = Qdp/dp with dp = dA*ps0 + dB*ps_v
  const int tracer = test == 11 ? 0 : 1;
  const size_t estride = (size_t)16 * 72 * cfg.qsize_d * 2;
  std::vector<double> qdp((size_t)mesh.nelem * estride), ps((size_t)mesh.nelem * 16);
  std::vector<double> q_i, q_f;
  auto mixing_ratio = [&](int tl, std::vector<double>& out) {
    TSE_CALL(tse_copy_qdp_d2h(h, qdp.data(), (long long)estride, tl));
    TSE_CALL(tse_get_dp3d_ps(h, nullptr, 0, ps.data(), 16));
    out.assign((size_t)mesh.nelem * 72 * 16, 0.0);
    for (int e = 0; e < mesh.nelem; ++e)
      for (int k = 0; k < 72; ++k)
        for (int n = 0; n < 16; ++n) {
          const double dp = (hv.hyai[k + 1] - hv.hyai[k]) * hv.ps0 + (hv.hybi[k + 1] - hv.hybi[k]) * ps[(size_t)e * 16 + n];
          out[((size_t)e * 72 + k) * 16 + n] = qdp[(size_t)e * estride + (size_t)(tl - 1) * 16 * 72 * cfg.qsize_d + ((size_t)tracer * 72 + k) * 16 + n] / dp;
        }
  };
  if (tracer < Q) mixing_ratio(1, q_i);

  // ---- main time-stepping loop (prim_main.F90:142-175) ----
  std::printf(" Entering main timestepping loop\n");
  int nstep = 0;
  TSE_CALL(tse_synchronize(h));
  const auto t0 = std::chrono::steady_clock::now();
  while (nstep < nEndStep) {
    const int before = nstep;
    TSE_CALL(tse_prim_run_subcycle(h, tstep, &nstep));
    if (nstep / statefreq != before / statefreq || nstep >= nEndStep) printstate(nstep, nstep % 2 == 0 ? 1 : 2);
  }
  TSE_CALL(tse_synchronize(h));
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf(" Finished main timestepping loop %d\n", nstep);
  std::printf(" prim_run wall time %.3f s  (%.2f tracer-steps/s, %.4f model-days/wall-s; GPU timers: prim_run %.1f ms, euler_step %.1f ms, vertical_remap %.1f ms)\n",
              wall, Q * (double)nstep / wall, nstep * tstep / 86400.0 / wall, tse_timer_ms(h, "prim_run"), tse_timer_ms(h, "euler_step"),
              tse_timer_ms(h, "vertical_remap"));

  // ---- error norms (test/dcmip1-1/dcmip1-1_error_norm_ng.ncl:41-77, dcmip1-2_error_norm_ng.ncl:42-77) on the unique columns ----
  if (tracer < Q) {
    mixing_ratio(nstep % 2 == 0 ? 1 : 2, q_f);
    const float pi32 = std::acos(-1.0f);
    const double rad = (double)(pi32 / 180.0f), R = (double)6.37122e6f, dlat = 0.5 * (double)pi32 / (cfg.ne * 3);
    const double H = 287.04 * 300.0 / 9.80616;
    std::vector<double> dh(72);
    double base = 0.0;
    for (int i = 1; i <= 72; ++i) {
      const int k = 72 - i;
      const double z = H * std::log(1.0 / (hv.hyam[k] + hv.hybm[k]));
      dh[k] = 2.0 * (z - base);
      base += dh[k];
    }
    // unique-point ownership: lowest global id among the elements sharing a node (dof_mod.F90:42-59,322-357)
    auto owns = [&](int e, int n) {
      const int i = n & 3, j = n >> 2;
      auto lose = [&](int d) { const int b = mesh.nbr[(size_t)e * 8 + d]; return b >= 0 && b < e; };
      if (j == 0 && lose(tse::SOUTH)) return false;
      if (i == 3 && lose(tse::EAST)) return false;
      if (j == 3 && lose(tse::NORTH)) return false;
      if (i == 0 && lose(tse::WEST)) return false;
      if (n == 0 && lose(tse::SWEST)) return false;
      if (n == 3 && lose(tse::SEAST)) return false;
      if (n == 12 && lose(tse::NWEST)) return false;
      if (n == 15 && lose(tse::NEAST)) return false;
      return true;
    };
    long double sum_qi = 0;
    size_t cnt = 0;
    for (int e = 0; e < mesh.nelem; ++e)
      for (int n = 0; n < 16; ++n)
        if (owns(e, n))
          for (int k = 0; k < 72; ++k) {
            sum_qi += q_i[((size_t)e * 72 + k) * 16 + n];
            ++cnt;
          }
    const double mean_qi = (double)(sum_qi / cnt);
    long double n1 = 0, d1 = 0, n2 = 0, d2 = 0;
    double ninf = 0, dinf = 0, qmax = -1e300, qmin = 1e300;
    for (int e = 0; e < mesh.nelem; ++e)
      for (int n = 0; n < 16; ++n) {
        if (!owns(e, n)) continue;
        const double lat = mesh.lat[(size_t)e * 16 + n] * (180.0 / M_PI) * rad;
        for (int k = 0; k < 72; ++k) {
          const double dV = (R * std::cos(lat) * dlat) * (R * dlat) * dh[k];
          const size_t i = ((size_t)e * 72 + k) * 16 + n;
          const double dq = q_f[i] - q_i[i], dev = std::fabs(q_i[i] - mean_qi);
          n1 += std::fabs(dq) * dV; d1 += dev * dV;
          n2 += dq * dq * dV; d2 += dev * dev * dV;
          ninf = std::max(ninf, std::fabs(dq) * dV); dinf = std::max(dinf, dev * dV);
          qmax = std::max(qmax, q_f[i]); qmin = std::min(qmin, q_f[i]);
        }
      }
    std::printf(" %s  L1 = %.7f  L2 = %.7f  Linf = %.7f  q_max = %.7f  q_min = %.6e\n", test_case.c_str(), (double)(n1 / d1),
                (double)std::sqrt((double)(n2 / d2)), ninf / dinf, qmax, qmin);
  }
  TSE_CALL(tse_finalize(h));
  return 0;
}
