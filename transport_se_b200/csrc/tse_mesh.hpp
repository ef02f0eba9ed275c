// Host-side cubed-sphere mesh for the transport_se tracer-advection path.
//
// This is the "caller side" of the hot path: it produces exactly the inputs
// the reference's Fortran host hands to the advection code -- per-element
// metric terms (element_t: spheremp, rspheremp, metdet, rmetdet, Dinv, spherep,
// reference src/share/element_mod.F90:112-221), the derivative matrix
// (derivative_t%Dvv, src/share/derivative_mod.F90:451-486) and the edge
// descriptors (EdgeDescriptor_t putmapP/getmapP/reverse,
// src/share/edge_mod.F90:31-43) plus the per-neighbour-rank exchange cycles
// (Schedule_t/Cycle_t, src/share/schedtype_mod.F90:7-29).
//
// It is NOT a translation of cube_mod's CubeTopology tables: connectivity is
// derived from exact integer vertex coordinates on the cube surface.
#pragma once
#include <cstdint>
#include <vector>

namespace tse {

constexpr int NP = 4;
constexpr int NPSQ = 16;
// direction indices, 0-based version of control_mod.F90:173-181
enum Dir { WEST = 0, EAST = 1, SOUTH = 2, NORTH = 3, SWEST = 4, SEAST = 5, NWEST = 6, NEAST = 7 };

struct GLL {
  double pts[NP];   // rounded from quad precision
  double wts[NP];
  double dvv[NPSQ]; // Fortran Dvv(i,l) stored at [i + 4*l]
  double mp[NPSQ];  // w_i*w_j at [i + 4*j]
};
const GLL& gll();

struct Mesh {
  int ne = 0;
  int nelem = 0;
  double alpha = 1.0;  // area correction (prim_driver_mod.F90:259-283)
  // element (0-based gid) -> face (1..6), ie, je (0-based)
  std::vector<int> face, ie, je;
  // topology, per element per direction
  std::vector<int> nbr;      // [e*8+d] neighbour gid or -1
  std::vector<int> nbr_dir;  // [e*8+d] direction index on the neighbour that faces back
  std::vector<uint8_t> rev;  // [e*8+d] reverse flag (edges only)
  // geometry, node index n = i + 4*j (i fastest, as Fortran (np,np))
  std::vector<double> lat, lon;                              // [e*16+n]
  std::vector<double> D, Dinv;                               // [(e*16+n)*4 + a + 2*b] = Fortran D(a,b,i,j)
  std::vector<double> metdet, rmetdet, spheremp, rspheremp;  // [e*16+n]
  // space-filling-curve index per element (cube_mod.F90:1501-1633)
  std::vector<int> sfc;

  explicit Mesh(int ne);
  // ordered DSS gather list for node n of element e, reference summation
  // order S,E,N,W then SW,SE,NE,NW (edge_mod.F90:648-742). Returns count (<=3).
  int gather(int e, int n, int src_elem[3], int src_node[3]) const;
};

// node on edge `d` (WEST..NORTH) at parameter t=0..3, and corner node of SWEST..NEAST
inline int edge_node(int d, int t) {
  switch (d) {
    case SOUTH: return t;
    case EAST: return 3 + 4 * t;
    case NORTH: return t + 12;
    default: return 4 * t;  // WEST
  }
}
inline int corner_node(int d) {
  switch (d) {
    case SWEST: return 0;
    case SEAST: return 3;
    case NWEST: return 12;
    default: return 15;  // NEAST
  }
}

// Per-rank view in the reference's own format.
struct LocalView {
  int rank = 0, nranks = 1;
  int nelemd = 0;
  std::vector<int> gid;      // local -> global (ascending, metagraph_mod.F90:317-323)
  std::vector<int> putmap;   // [le*8+d] 0-based buffer offset or -1
  std::vector<int> getmap;   // [le*8+d]
  std::vector<int> reverse;  // [le*8+d] 0/1
  int nbuf = 0;              // horizontal size of the edge buffer
  std::vector<int> cyc_rank, cyc_ptr, cyc_len;  // exchange cycles (one per neighbour rank)
};

// contiguous SFC chunks, first mod(nelem,nparts) chunks one larger (spacecurve_mod.F90:1218-1273)
std::vector<int> sfc_partition(const Mesh& m, int nparts);
LocalView make_local_view(const Mesh& m, const std::vector<int>& owner, int rank, int nranks);

}  // namespace tse
