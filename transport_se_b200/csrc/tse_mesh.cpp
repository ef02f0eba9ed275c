// Cubed-sphere mesh, metric terms, DSS connectivity, SFC partition (host only).
// See tse_mesh.hpp for scope; reference citations are relative to /root/reference.
#include "tse_mesh.hpp"

#include <quadmath.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <tuple>
#include <unordered_map>

namespace tse {

typedef __float128 quad;

// ---------------------------------------------------------------------------
// GLL points / weights / derivative matrix for np=4.
// The reference finds the points by Newton iteration in longdouble_kind=16
// (quadrature_mod.F90:305-470) and only then rounds to real_kind; the closed
// forms (+-1, +-1/sqrt(5); 1/6, 5/6) evaluated in binary128 round to the same
// doubles.  Dvv follows dvvinit (derivative_mod.F90:451-486) in binary128.
// ---------------------------------------------------------------------------
static quad gll_q(int i) {
  static const quad r5 = 1.0Q / sqrtq(5.0Q);
  switch (i) {
    case 0: return -1.0Q;
    case 1: return -r5;
    case 2: return r5;
    default: return 1.0Q;
  }
}
static quad gllw_q(int i) { return (i == 0 || i == 3) ? 1.0Q / 6.0Q : 5.0Q / 6.0Q; }

// Legendre P_0..P_n by the three-term recurrence (quadrature_mod.F90:725-748)
static void legendre_q(quad x, int n, quad* leg) {
  quad p3 = 1.0Q, p2, p1;
  leg[0] = p3;
  if (n != 0) {
    p2 = p3;
    p3 = x;
    leg[1] = p3;
    for (int k = 2; k <= n; ++k) {
      p1 = p2;
      p2 = p3;
      p3 = ((2 * k - 1) * x * p2 - (k - 1) * p1) / k;
      leg[k] = p3;
    }
  }
}

const GLL& gll() {
  static GLL g;
  static bool init = false;
  if (!init) {
    quad leg[NP][NP];  // leg[i][k] = P_k(x_i)
    for (int i = 0; i < NP; ++i) legendre_q(gll_q(i), NP - 1, leg[i]);
    quad dvv[NP][NP];  // dvv[j][i] as Fortran dvv(j,i)
    for (int j = 0; j < NP; ++j)
      for (int i = 0; i < NP; ++i)
        dvv[j][i] = (i == j) ? 0.0Q : (1.0Q / (gll_q(i) - gll_q(j))) * leg[i][NP - 1] / leg[j][NP - 1];
    dvv[NP - 1][NP - 1] = quad(NP * (NP - 1)) / 4.0Q;
    dvv[0][0] = -quad(NP * (NP - 1)) / 4.0Q;
    for (int i = 0; i < NP; ++i) {
      g.pts[i] = (double)gll_q(i);
      g.wts[i] = (double)gllw_q(i);
    }
    // deriv%Dvv(a,b) = dvv(a,b): column-major -> flat [a + 4*b]
    for (int a = 0; a < NP; ++a)
      for (int b = 0; b < NP; ++b) g.dvv[a + 4 * b] = (double)dvv[a][b];
    for (int j = 0; j < NP; ++j)
      for (int i = 0; i < NP; ++i) g.mp[i + 4 * j] = (double)(gllw_q(i) * gllw_q(j));  // mass_matrix_mod.F90:63
    init = true;
  }
  return g;
}

// ---------------------------------------------------------------------------
// Space-filling curve (spacecurve_mod.F90). Each generator visits the k x k
// sub-cells of its level in a fixed order; for sub-cell s it recurses with
// (main axis, main dir, joiner axis, joiner dir) derived from its own by the
// rule {lma,lmd,lja,ljd}: lma 0=ma 1=other axis; lmd +-1 = +-md;
// lja 0=ma 1=other 2=inherit ja; ljd +-1 = +-md, 0 = inherit jd.
// Tables restate hilbert (:683-769), PeanoM (:505-681), Cinco (:39-503).
// ---------------------------------------------------------------------------
namespace {
struct Rule { int8_t lma, lmd, lja, ljd; };
const Rule kHilbert[4] = {{1, 1, 1, 1}, {0, 1, 0, 1}, {0, 1, 1, -1}, {1, -1, 2, 0}};
const Rule kPeano[9] = {{1, 1, 1, 1}, {1, 1, 1, 1}, {0, 1, 0, 1}, {0, 1, 0, 1}, {0, 1, 1, -1},
                        {0, -1, 0, -1}, {1, -1, 1, -1}, {1, -1, 0, 1}, {0, 1, 2, 0}};
const Rule kCinco[25] = {{0, 1, 0, 1}, {0, 1, 0, 1}, {1, 1, 1, 1}, {1, 1, 1, 1}, {1, 1, 0, -1},
                         {1, -1, 1, -1}, {0, -1, 0, -1}, {0, -1, 1, 1}, {1, 1, 1, 1}, {1, 1, 1, 1},
                         {0, 1, 0, 1}, {0, 1, 1, -1}, {1, -1, 0, 1}, {1, 1, 1, 1}, {0, 1, 0, 1},
                         {0, 1, 0, 1}, {0, 1, 1, -1}, {0, -1, 0, -1}, {1, -1, 1, -1}, {1, -1, 0, 1},
                         {0, 1, 1, -1}, {0, -1, 0, -1}, {1, -1, 1, -1}, {1, -1, 0, 1}, {0, 1, 2, 0}};

struct Curve {
  int n;
  std::vector<int> ordered;  // [x + n*y]
  std::vector<int> factors;  // as Factor(): all 2s, then 3s, then 5s
  int pos[2] = {0, 0};
  int vcnt = 0;
  void gen(int l, int type, int ma, int md, int ja, int jd) {
    const Rule* r = type == 2 ? kHilbert : type == 3 ? kPeano : kCinco;
    const int cnt = type * type;
    const int ltype = l > 1 ? factors[l - 2] : 0;
    for (int s = 0; s < cnt; ++s) {
      int lma = r[s].lma ? (ma + 1) % 2 : ma;
      int lmd = r[s].lmd * md;
      int lja = r[s].lja == 2 ? ja : (r[s].lja ? (ma + 1) % 2 : ma);
      int ljd = r[s].ljd == 0 ? jd : r[s].ljd * md;
      if (l > 1) {
        gen(l - 1, ltype, lma, lmd, lja, ljd);
      } else {  // IncrementCurve (:771-784)
        ordered[pos[0] + n * pos[1]] = vcnt++;
        pos[lja] += ljd;
      }
    }
  }
};

bool factor235(int num, std::vector<int>& f) {
  f.clear();
  int t = num;
  for (int p : {2, 3, 5})
    while (t % p == 0) { f.push_back(p); t /= p; }
  return t == 1 && !f.empty();
}

// Mesh(i,j) visitation order for an n x n face, 0-based [i + n*j] (GenSpaceCurve :1011-1040)
std::vector<int> gen_space_curve(int n) {
  Curve c;
  c.n = n;
  c.ordered.assign((size_t)n * n, 0);
  if (n == 1) return c.ordered;
  if (!factor235(n, c.factors)) throw std::runtime_error("gen_space_curve: size not 2^a 3^b 5^c");
  int level = (int)c.factors.size();
  c.gen(level, c.factors[level - 1], 0, 1, 0, 1);  // map(): GenCurve(l,type,0,1,0,1)
  return c.ordered;
}

// face curve for arbitrary ne (cube_mod.F90:1500-1571): factorable -> direct,
// otherwise project the curve of the next power of two.
std::vector<int> face_curve(int ne) {
  std::vector<int> f;
  if (ne == 1 || factor235(ne, f)) return gen_space_curve(ne);
  int ne2 = 1;
  while (ne2 < ne) ne2 *= 2;
  std::vector<int> mesh2 = gen_space_curve(ne2);
  std::vector<int> map_i((size_t)ne2 * ne2, 0), map_j((size_t)ne2 * ne2, 0);
  for (int j = 1; j <= ne; ++j)
    for (int i = 1; i <= ne; ++i) {
      int i2 = (int)std::lround(((i - .5) / ne) * ne2 + .5);
      int j2 = (int)std::lround(((j - .5) / ne) * ne2 + .5);
      i2 = std::min(std::max(i2, 1), ne2);
      j2 = std::min(std::max(j2, 1), ne2);
      map_i[(i2 - 1) + ne2 * (j2 - 1)] = i;
      map_j[(i2 - 1) + ne2 * (j2 - 1)] = j;
    }
  std::vector<int> where((size_t)ne2 * ne2);
  for (int idx = 0; idx < ne2 * ne2; ++idx) where[mesh2[idx]] = idx;
  std::vector<int> mesh((size_t)ne * ne, 0);
  int sfc_index = 0;
  for (int k = 0; k < ne2 * ne2; ++k) {
    int idx = where[k];
    if (map_i[idx] != 0) mesh[(map_i[idx] - 1) + ne * (map_j[idx] - 1)] = sfc_index++;
  }
  return mesh;
}
}  // namespace

// ---------------------------------------------------------------------------
// Geometry of one element (cube_mod.F90: set_corner_coordinates :1280-1330,
// coordinates_atomic :142-181, metric_atomic :241-486, dmap_equiangular
// :582-641, vmap :658-743; coordinate_systems_mod.F90:308-402,508-525).
// ---------------------------------------------------------------------------
namespace {
const double DD_PI = 3.141592653589793238462643383279;
const double DIST_THRESHOLD = 1.0e-9;

void projectpoint(double cx, double cy, int face, double& lon, double& lat) {
  double x = std::tan(cx), y = std::tan(cy);
  double r = std::sqrt(1.0 + x * x + y * y);
  switch (face) {
    case 1: lat = std::asin(y / r); lon = std::atan2(x, 1.0); break;
    case 2: lat = std::asin(y / r); lon = std::atan2(1.0, -x); break;
    case 3: lat = std::asin(y / r); lon = std::atan2(-x, -1.0); break;
    case 4: lat = std::asin(y / r); lon = std::atan2(-1.0, x); break;
    case 5:
      lon = (std::fabs(y) > DIST_THRESHOLD || std::fabs(x) > DIST_THRESHOLD) ? std::atan2(x, y) : 0.0;
      lat = std::asin(-1.0 / r);
      break;
    default:
      lon = (std::fabs(y) > DIST_THRESHOLD || std::fabs(x) > DIST_THRESHOLD) ? std::atan2(x, -y) : 0.0;
      lat = std::asin(1.0 / r);
      break;
  }
  if (lon < 0.0) lon += 2.0 * DD_PI;
}

void vmap(double D[4], double x1, double x2, int face) {  // D[a+2b] = D(a,b)
  double r = std::sqrt(1.0 + std::tan(x1) * std::tan(x1) + std::tan(x2) * std::tan(x2));
  double D11, D12, D21, D22;
  if (face >= 1 && face <= 4) {
    D11 = 1.0 / (r * std::cos(x1));
    D12 = 0.0;
    D21 = -std::tan(x1) * std::tan(x2) / (std::cos(x1) * r * r);
    D22 = 1.0 / (r * r * std::cos(x1) * std::cos(x2) * std::cos(x2));
  } else {
    double poledist = std::sqrt(std::tan(x1) * std::tan(x1) + std::tan(x2) * std::tan(x2));
    if (poledist <= DIST_THRESHOLD) {
      D11 = 1.0; D12 = 0.0; D21 = 0.0; D22 = 1.0;
    } else if (face == 6) {
      D11 = -std::tan(x2) / (poledist * std::cos(x1) * std::cos(x1) * r);
      D12 = std::tan(x1) / (poledist * std::cos(x2) * std::cos(x2) * r);
      D21 = -std::tan(x1) / (poledist * std::cos(x1) * std::cos(x1) * r * r);
      D22 = -std::tan(x2) / (poledist * std::cos(x2) * std::cos(x2) * r * r);
    } else {
      D11 = std::tan(x2) / (poledist * std::cos(x1) * std::cos(x1) * r);
      D12 = -std::tan(x1) / (poledist * std::cos(x2) * std::cos(x2) * r);
      D21 = std::tan(x1) / (poledist * std::cos(x1) * std::cos(x1) * r * r);
      D22 = std::tan(x2) / (poledist * std::cos(x2) * std::cos(x2) * r * r);
    }
  }
  D[0] = D11; D[2] = D12; D[1] = D21; D[3] = D22;
}

struct ElemGeom {
  double lat[NPSQ], lon[NPSQ], D[NPSQ * 4], Dinv[NPSQ * 4], metdet[NPSQ], rmetdet[NPSQ];
};

void element_geometry(int ne, int face, int ie, int je, double alpha, ElemGeom& g) {
  const double xs = -0.25 * DD_PI, xe = 0.25 * DD_PI;
  const double dx = (xe - xs) / ne, dy = (xe - xs) / ne;
  const double startx = xs + ie * dx, starty = xs + je * dy;
  const double cx[4] = {startx, startx + dx, startx + dx, startx};
  const double cy[4] = {starty, starty, starty + dy, starty + dy};

  // cartp = element_var_coordinates (element_mod.F90:284-310), quad p/q
  double cartx[NPSQ], carty[NPSQ];
  quad p[NP], q[NP];
  for (int i = 0; i < NP; ++i) {
    p[i] = (1.0Q - gll_q(i)) / 2.0Q;
    q[i] = (1.0Q + gll_q(i)) / 2.0Q;
  }
  for (int j = 0; j < NP; ++j)
    for (int i = 0; i < NP; ++i) {
      cartx[i + 4 * j] = (double)(p[i] * p[j] * cx[0] + q[i] * p[j] * cx[1] + q[i] * q[j] * cx[2] + p[i] * q[j] * cx[3]);
      carty[i + 4 * j] = (double)(p[i] * p[j] * cy[0] + q[i] * p[j] * cy[1] + q[i] * q[j] * cy[2] + p[i] * q[j] * cy[3]);
    }
  // u2qmap = elem_jacobians (cube_mod.F90:190-207)
  const int n11 = 0, n41 = 3, n44 = 15, n14 = 12;
  double u2q[4][2];
  u2q[0][0] = (cartx[n11] + cartx[n41] + cartx[n44] + cartx[n14]) / 4.0;
  u2q[0][1] = (carty[n11] + carty[n41] + carty[n44] + carty[n14]) / 4.0;
  u2q[1][0] = (-cartx[n11] + cartx[n41] + cartx[n44] - cartx[n14]) / 4.0;
  u2q[1][1] = (-carty[n11] + carty[n41] + carty[n44] - carty[n14]) / 4.0;
  u2q[2][0] = (-cartx[n11] - cartx[n41] + cartx[n44] + cartx[n14]) / 4.0;
  u2q[2][1] = (-carty[n11] - carty[n41] + carty[n44] + carty[n14]) / 4.0;
  u2q[3][0] = (cartx[n11] - cartx[n41] + cartx[n44] - cartx[n14]) / 4.0;
  u2q[3][1] = (carty[n11] - carty[n41] + carty[n44] - carty[n14]) / 4.0;

  const GLL& G = gll();
  for (int j = 0; j < NP; ++j)
    for (int i = 0; i < NP; ++i) {
      const int n = i + 4 * j;
      // spherep: ref2sphere_equiangular_longdouble (cube_mod.F90:2482-2511):
      // the bilinear weights are formed in quad and stored to real_kind locals.
      {
        double pi_ = (double)((1.0Q - gll_q(i)) / 2.0Q), pj_ = (double)((1.0Q - gll_q(j)) / 2.0Q);
        double qi_ = (double)((1.0Q + gll_q(i)) / 2.0Q), qj_ = (double)((1.0Q + gll_q(j)) / 2.0Q);
        double x = pi_ * pj_ * cx[0] + qi_ * pj_ * cx[1] + qi_ * qj_ * cx[2] + pi_ * qj_ * cx[3];
        double y = pi_ * pj_ * cy[0] + qi_ * pj_ * cy[1] + qi_ * qj_ * cy[2] + pi_ * qj_ * cy[3];
        projectpoint(x, y, face, g.lon[n], g.lat[n]);
      }
      // D: dmap_equiangular with a,b = real_kind(gll points)
      const double a = G.pts[i], b = G.pts[j];
      double Jp11 = u2q[1][0] + u2q[3][0] * b;
      double Jp12 = u2q[2][0] + u2q[3][0] * a;
      double Jp21 = u2q[1][1] + u2q[3][1] * b;
      double Jp22 = u2q[2][1] + u2q[3][1] * a;
      double pi_ = (1 - a) / 2, pj_ = (1 - b) / 2, qi_ = (1 + a) / 2, qj_ = (1 + b) / 2;
      double x1 = pi_ * pj_ * cx[0] + qi_ * pj_ * cx[1] + qi_ * qj_ * cx[2] + pi_ * qj_ * cx[3];
      double x2 = pi_ * pj_ * cy[0] + qi_ * pj_ * cy[1] + qi_ * qj_ * cy[2] + pi_ * qj_ * cy[3];
      double t[4];
      vmap(t, x1, x2, face);
      double* D = &g.D[n * 4];
      D[0] = t[0] * Jp11 + t[2] * Jp21;  // D(1,1)
      D[2] = t[0] * Jp12 + t[2] * Jp22;  // D(1,2)
      D[1] = t[1] * Jp11 + t[3] * Jp21;  // D(2,1)
      D[3] = t[1] * Jp12 + t[3] * Jp22;  // D(2,2)
      double detD = D[0] * D[3] - D[2] * D[1];
      double* Di = &g.Dinv[n * 4];
      Di[0] = D[3] / detD;   // Dinv(1,1)
      Di[2] = -D[2] / detD;  // Dinv(1,2)
      Di[1] = -D[1] / detD;  // Dinv(2,1)
      Di[3] = D[0] / detD;   // Dinv(2,2)
      g.metdet[n] = std::fabs(detD);
      g.rmetdet[n] = 1.0 / std::fabs(detD);
      // optional normalisation (cube_mod.F90:478-484)
      const double sa = std::sqrt(alpha);
      for (int c = 0; c < 4; ++c) {
        D[c] = D[c] * sa;
        Di[c] = Di[c] / sa;
      }
      g.metdet[n] = g.metdet[n] * alpha;
      g.rmetdet[n] = g.rmetdet[n] / alpha;
    }
}
}  // namespace

// ---------------------------------------------------------------------------
// Mesh
// ---------------------------------------------------------------------------
int Mesh::gather(int e, int n, int src_elem[3], int src_node[3]) const {
  int cnt = 0;
  const int i = n & 3, j = n >> 2;
  auto from_edge = [&](int d, int t) {
    int b = nbr[e * 8 + d];
    if (b < 0) return;
    int bd = nbr_dir[e * 8 + d];
    // the sender packs edge parameter t' into slot t'(or 3-t' when reversed); we read slot t
    int tb = rev[b * 8 + bd] ? 3 - t : t;
    src_elem[cnt] = b;
    src_node[cnt] = edge_node(bd, tb);
    ++cnt;
  };
  auto from_corner = [&](int d) {
    int b = nbr[e * 8 + d];
    if (b < 0) return;
    src_elem[cnt] = b;
    src_node[cnt] = corner_node(nbr_dir[e * 8 + d]);
    ++cnt;
  };
  // unpack order of edge_mod.F90:678-702: South, East, North, West, then corners
  if (j == 0) from_edge(SOUTH, i);
  if (i == 3) from_edge(EAST, j);
  if (j == 3) from_edge(NORTH, i);
  if (i == 0) from_edge(WEST, j);
  if (i == 0 && j == 0) from_corner(SWEST);
  if (i == 3 && j == 0) from_corner(SEAST);
  if (i == 3 && j == 3) from_corner(NEAST);
  if (i == 0 && j == 3) from_corner(NWEST);
  return cnt;
}

Mesh::Mesh(int ne_) : ne(ne_) {
  if (ne < 2) throw std::runtime_error("Mesh: ne must be >= 2");
  nelem = 6 * ne * ne;
  face.resize(nelem); ie.resize(nelem); je.resize(nelem);
  for (int e = 0; e < nelem; ++e) {  // convert_gbl_index (cube_mod.F90:1418-1430)
    face[e] = e / (ne * ne) + 1;
    ie[e] = e % ne;
    je[e] = e / ne - (face[e] - 1) * ne;
  }
  // --- topology from integer cube-surface coordinates of the element corners ---
  // Face-local lattice (u,v) in [-ne,ne] (step 2) is embedded the same way
  // unit_face_based_cube_to_unit_sphere orients the faces.
  auto vertex3 = [&](int f, int iu, int iv) -> std::array<int, 3> {
    int u = 2 * iu - ne, v = 2 * iv - ne, c = ne;
    switch (f) {
      case 1: return {c, u, v};
      case 2: return {-u, c, v};
      case 3: return {-c, -u, v};
      case 4: return {u, -c, v};
      case 5: return {v, u, -c};
      default: return {-v, u, c};
    }
  };
  std::map<std::array<int, 3>, int> vid;
  std::vector<std::array<int, 4>> ev(nelem);  // corner vertex ids: SW, SE, NE, NW
  for (int e = 0; e < nelem; ++e) {
    const int ci[4] = {0, 1, 1, 0}, cj[4] = {0, 0, 1, 1};
    for (int c = 0; c < 4; ++c) {
      auto key = vertex3(face[e], ie[e] + ci[c], je[e] + cj[c]);
      auto it = vid.find(key);
      if (it == vid.end()) it = vid.emplace(key, (int)vid.size()).first;
      ev[e][c] = it->second;
    }
  }
  if ((int)vid.size() != 6 * ne * ne + 2) throw std::runtime_error("Mesh: Euler test failed");  // cube_mod.F90:1403
  std::vector<std::vector<int>> v2e(vid.size());
  for (int e = 0; e < nelem; ++e)
    for (int c = 0; c < 4; ++c) v2e[ev[e][c]].push_back(e);

  nbr.assign((size_t)nelem * 8, -1);
  nbr_dir.assign((size_t)nelem * 8, -1);
  rev.assign((size_t)nelem * 8, 0);
  // edge d -> (start corner, end corner) along increasing local coordinate
  auto edge_corners = [](int d, int& a, int& b) {
    switch (d) {
      case SOUTH: a = 0; b = 1; break;
      case EAST: a = 1; b = 2; break;
      case NORTH: a = 3; b = 2; break;
      default: a = 0; b = 3; break;  // WEST
    }
  };
  auto corner_of_dir = [](int d) { return d == SWEST ? 0 : d == SEAST ? 1 : d == NEAST ? 2 : 3; };
  const int corner_dirs[4] = {SWEST, SEAST, NEAST, NWEST};
  for (int e = 0; e < nelem; ++e) {
    for (int d = 0; d < 4; ++d) {
      int a, b;
      edge_corners(d, a, b);
      int va = ev[e][a], vb = ev[e][b];
      for (int o : v2e[va]) {
        if (o == e) continue;
        for (int od = 0; od < 4; ++od) {
          int oa, ob;
          edge_corners(od, oa, ob);
          int wa = ev[o][oa], wb = ev[o][ob];
          if ((wa == va && wb == vb) || (wa == vb && wb == va)) {
            nbr[e * 8 + d] = o;
            nbr_dir[e * 8 + d] = od;
            rev[e * 8 + d] = (wa != va);
          }
        }
      }
      if (nbr[e * 8 + d] < 0) throw std::runtime_error("Mesh: missing edge neighbour");
    }
    for (int c = 0; c < 4; ++c) {
      int d = corner_dirs[c];
      int v = ev[e][c];
      for (int o : v2e[v]) {
        if (o == e) continue;
        bool is_edge_nbr = false;
        for (int dd = 0; dd < 4; ++dd)
          if (nbr[e * 8 + dd] == o) is_edge_nbr = true;
        if (is_edge_nbr) continue;
        int oc = -1;
        for (int k = 0; k < 4; ++k)
          if (ev[o][k] == v) oc = k;
        nbr[e * 8 + d] = o;
        nbr_dir[e * 8 + d] = corner_dirs[oc];
      }
    }
    (void)corner_of_dir;
  }

  // --- space-filling curve (cube_mod.F90:1574-1633): faces stitched 1,2,6,4,5,3 ---
  {
    std::vector<int> M = face_curve(ne);  // M[i + ne*j], 0-based i,j
    auto Mat = [&](int i, int j) { return M[(i - 1) + ne * (j - 1)]; };  // 1-based
    sfc.resize(nelem);
    const int n2 = ne * ne;
    for (int j = 1; j <= ne; ++j)
      for (int i = 1; i <= ne; ++i) {
        auto id = [&](int f) { return (i - 1) + ne * (j - 1) + n2 * (f - 1); };
        sfc[id(1)] = 0 * n2 + Mat(i, ne - j + 1);
        sfc[id(2)] = 1 * n2 + Mat(i, ne - j + 1);
        sfc[id(6)] = 2 * n2 + Mat(ne - i + 1, ne - j + 1);
        sfc[id(4)] = 3 * n2 + Mat(ne - j + 1, i);
        sfc[id(5)] = 4 * n2 + Mat(i, j);
        sfc[id(3)] = 5 * n2 + Mat(i, j);
      }
  }

  // --- geometry: two passes for the area correction (prim_driver_mod.F90:259-283) ---
  lat.resize((size_t)nelem * 16); lon.resize((size_t)nelem * 16);
  D.resize((size_t)nelem * 64); Dinv.resize((size_t)nelem * 64);
  metdet.resize((size_t)nelem * 16); rmetdet.resize((size_t)nelem * 16);
  spheremp.resize((size_t)nelem * 16); rspheremp.resize((size_t)nelem * 16);
  const GLL& G = gll();
  auto fill = [&](double a) {
#pragma omp parallel for schedule(static)
    for (int e = 0; e < nelem; ++e) {
      ElemGeom g;
      element_geometry(ne, face[e], ie[e], je[e], a, g);
      std::memcpy(&lat[(size_t)e * 16], g.lat, sizeof g.lat);
      std::memcpy(&lon[(size_t)e * 16], g.lon, sizeof g.lon);
      std::memcpy(&D[(size_t)e * 64], g.D, sizeof g.D);
      std::memcpy(&Dinv[(size_t)e * 64], g.Dinv, sizeof g.Dinv);
      std::memcpy(&metdet[(size_t)e * 16], g.metdet, sizeof g.metdet);
      std::memcpy(&rmetdet[(size_t)e * 16], g.rmetdet, sizeof g.rmetdet);
    }
  };
  fill(1.0);
  {
    // aratio(ie) = sum(mp*metdet); global sum through repro_sum (order-free
    // fixed-point); binary128 accumulation gives the same correctly rounded sum.
    quad area = 0;
    for (int e = 0; e < nelem; ++e) {
      double s = 0;
      for (int n = 0; n < 16; ++n) s += G.mp[n] * metdet[(size_t)e * 16 + n];
      area += s;
    }
    alpha = 4.0 * DD_PI / (double)area;
  }
  fill(alpha);
  // mass matrix (mass_matrix_mod.F90:97-116): spheremp = mp*metdet, rspheremp = 1/DSS(spheremp)
  for (size_t k = 0; k < (size_t)nelem * 16; ++k) spheremp[k] = G.mp[k & 15] * metdet[k];
#pragma omp parallel for schedule(static)
  for (int e = 0; e < nelem; ++e)
    for (int n = 0; n < 16; ++n) {
      int se[3], sn[3];
      int c = gather(e, n, se, sn);
      double v = spheremp[(size_t)e * 16 + n];
      for (int k = 0; k < c; ++k) v = v + spheremp[(size_t)se[k] * 16 + sn[k]];
      rspheremp[(size_t)e * 16 + n] = 1.0 / v;
    }
}

std::vector<int> sfc_partition(const Mesh& m, int nparts) {
  // genspacepart (spacecurve_mod.F90:1218-1273)
  std::vector<int> owner(m.nelem);
  const int nelemd = m.nelem / nparts;
  const int extra = m.nelem % nparts;
  const int s1 = extra * (nelemd + 1);
  for (int e = 0; e < m.nelem; ++e) {
    int id = m.sfc[e];
    if (id <= s1 && extra > 0) {
      owner[e] = std::min(id / (nelemd + 1), nparts - 1);
    } else {
      id -= s1;
      owner[e] = extra + id / nelemd;
    }
    if (owner[e] >= nparts) owner[e] = nparts - 1;
  }
  return owner;
}

LocalView make_local_view(const Mesh& m, const std::vector<int>& owner, int rank, int nranks) {
  LocalView v;
  v.rank = rank;
  v.nranks = nranks;
  std::vector<int> g2l(m.nelem, -1);
  for (int e = 0; e < m.nelem; ++e)
    if (owner[e] == rank) {
      g2l[e] = (int)v.gid.size();
      v.gid.push_back(e);
    }
  v.nelemd = (int)v.gid.size();
  v.putmap.assign((size_t)v.nelemd * 8, -1);
  v.getmap.assign((size_t)v.nelemd * 8, -1);
  v.reverse.assign((size_t)v.nelemd * 8, 0);
  int next = 0;
  // intra-rank messages: sender's putmap == receiver's getmap (schedule_mod.F90:150-151)
  for (int le = 0; le < v.nelemd; ++le) {
    int e = v.gid[le];
    for (int d = 0; d < 8; ++d) {
      v.reverse[le * 8 + d] = m.rev[e * 8 + d];
      int b = m.nbr[e * 8 + d];
      if (b < 0 || owner[b] != rank) continue;
      int len = d < 4 ? NP : 1;
      v.putmap[le * 8 + d] = next;
      v.getmap[g2l[b] * 8 + m.nbr_dir[e * 8 + d]] = next;
      next += len;
    }
  }
  // inter-rank: one slab per neighbour rank; both sides order the shared
  // (element,direction) pairs by the same key so offsets agree.
  struct Item { int lo_gid, lo_dir, hi_gid, hi_dir, le, d, len; };
  std::map<int, std::vector<Item>> by_rank;
  for (int le = 0; le < v.nelemd; ++le) {
    int e = v.gid[le];
    for (int d = 0; d < 8; ++d) {
      int b = m.nbr[e * 8 + d];
      if (b < 0 || owner[b] == rank) continue;
      int bd = m.nbr_dir[e * 8 + d];
      Item it;
      if (e < b) { it.lo_gid = e; it.lo_dir = d; it.hi_gid = b; it.hi_dir = bd; }
      else { it.lo_gid = b; it.lo_dir = bd; it.hi_gid = e; it.hi_dir = d; }
      it.le = le; it.d = d; it.len = d < 4 ? NP : 1;
      by_rank[owner[b]].push_back(it);
    }
  }
  for (auto& kv : by_rank) {
    auto& items = kv.second;
    std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) {
      return std::tie(a.lo_gid, a.lo_dir, a.hi_gid, a.hi_dir) < std::tie(b.lo_gid, b.lo_dir, b.hi_gid, b.hi_dir);
    });
    v.cyc_rank.push_back(kv.first);
    v.cyc_ptr.push_back(next);
    int start = next;
    for (auto& it : items) {
      v.putmap[it.le * 8 + it.d] = next;
      v.getmap[it.le * 8 + it.d] = next;
      next += it.len;
    }
    v.cyc_len.push_back(next - start);
  }
  v.nbuf = next;
  return v;
}

}  // namespace tse

// ---------------------------------------------------------------------------
// C API (ctypes-friendly) for tests, the Python host mirror and the drivers.
// ---------------------------------------------------------------------------
extern "C" {

void tse_gll(double* pts4, double* wts4, double* dvv16, double* mp16) {
  const tse::GLL& g = tse::gll();
  if (pts4) std::memcpy(pts4, g.pts, sizeof g.pts);
  if (wts4) std::memcpy(wts4, g.wts, sizeof g.wts);
  if (dvv16) std::memcpy(dvv16, g.dvv, sizeof g.dvv);
  if (mp16) std::memcpy(mp16, g.mp, sizeof g.mp);
}

void* tse_mesh_create(int ne) {
  try {
    return new tse::Mesh(ne);
  } catch (const std::exception& ex) {
    std::fprintf(stderr, "tse_mesh_create: %s\n", ex.what());
    return nullptr;
  }
}
void tse_mesh_destroy(void* m) { delete (tse::Mesh*)m; }
int tse_mesh_nelem(void* m) { return ((tse::Mesh*)m)->nelem; }
double tse_mesh_alpha(void* m) { return ((tse::Mesh*)m)->alpha; }

// Copies global arrays; any pointer may be NULL.
void tse_mesh_get(void* mp, double* lat, double* lon, double* D, double* Dinv, double* metdet, double* rmetdet,
                  double* spheremp, double* rspheremp, int* nbr, int* nbr_dir, int* rev, int* sfc, int* face_ie_je) {
  auto* m = (tse::Mesh*)mp;
  size_t n16 = (size_t)m->nelem * 16, n64 = (size_t)m->nelem * 64, n8 = (size_t)m->nelem * 8;
  if (lat) std::memcpy(lat, m->lat.data(), n16 * 8);
  if (lon) std::memcpy(lon, m->lon.data(), n16 * 8);
  if (D) std::memcpy(D, m->D.data(), n64 * 8);
  if (Dinv) std::memcpy(Dinv, m->Dinv.data(), n64 * 8);
  if (metdet) std::memcpy(metdet, m->metdet.data(), n16 * 8);
  if (rmetdet) std::memcpy(rmetdet, m->rmetdet.data(), n16 * 8);
  if (spheremp) std::memcpy(spheremp, m->spheremp.data(), n16 * 8);
  if (rspheremp) std::memcpy(rspheremp, m->rspheremp.data(), n16 * 8);
  if (nbr) std::memcpy(nbr, m->nbr.data(), n8 * 4);
  if (nbr_dir) std::memcpy(nbr_dir, m->nbr_dir.data(), n8 * 4);
  if (rev)
    for (size_t k = 0; k < n8; ++k) rev[k] = m->rev[k];
  if (sfc) std::memcpy(sfc, m->sfc.data(), (size_t)m->nelem * 4);
  if (face_ie_je)
    for (int e = 0; e < m->nelem; ++e) {
      face_ie_je[3 * e] = m->face[e];
      face_ie_je[3 * e + 1] = m->ie[e];
      face_ie_je[3 * e + 2] = m->je[e];
    }
}

void tse_mesh_sfc_partition(void* mp, int nparts, int* owner) {
  auto* m = (tse::Mesh*)mp;
  std::vector<int> o = tse::sfc_partition(*m, nparts);
  std::memcpy(owner, o.data(), o.size() * 4);
}

// Local view: two-call protocol. First call with all outputs NULL returns sizes.
void* tse_local_view_create(void* mp, const int* owner, int rank, int nranks) {
  auto* m = (tse::Mesh*)mp;
  std::vector<int> o(owner, owner + m->nelem);
  return new tse::LocalView(tse::make_local_view(*m, o, rank, nranks));
}
void tse_local_view_destroy(void* v) { delete (tse::LocalView*)v; }
void tse_local_view_sizes(void* vp, int* nelemd, int* nbuf, int* ncycles) {
  auto* v = (tse::LocalView*)vp;
  *nelemd = v->nelemd;
  *nbuf = v->nbuf;
  *ncycles = (int)v->cyc_rank.size();
}
void tse_local_view_get(void* vp, int* gid, int* putmap, int* getmap, int* reverse, int* cyc_rank, int* cyc_ptr,
                        int* cyc_len) {
  auto* v = (tse::LocalView*)vp;
  if (gid) std::memcpy(gid, v->gid.data(), v->gid.size() * 4);
  if (putmap) std::memcpy(putmap, v->putmap.data(), v->putmap.size() * 4);
  if (getmap) std::memcpy(getmap, v->getmap.data(), v->getmap.size() * 4);
  if (reverse) std::memcpy(reverse, v->reverse.data(), v->reverse.size() * 4);
  if (cyc_rank) std::memcpy(cyc_rank, v->cyc_rank.data(), v->cyc_rank.size() * 4);
  if (cyc_ptr) std::memcpy(cyc_ptr, v->cyc_ptr.data(), v->cyc_ptr.size() * 4);
  if (cyc_len) std::memcpy(cyc_len, v->cyc_len.data(), v->cyc_len.size() * 4);
}

}  // extern "C"
