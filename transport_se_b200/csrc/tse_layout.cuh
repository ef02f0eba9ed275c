// Device data layout of the tracer-advection path.
//
// The reference stores every field inside element_t (AoS, element_mod.F90:20-78):
// Qdp(np,np,nlev,qsize_d,2) per element.  On the device the private copy is tiled
// for the access pattern of the stage kernels:
//
//   tracer field  Q[g][kc][q][el][kk][16]      g  = group of GE consecutive elements (internal order)
//   level field   F[g][kc][el][kk][16]         kc = chunk of KC=4 levels, kk = level in chunk
//   vn0           V[g][kc][c][el][kk][16]      el = element in group, 16 = (i,j) nodes, i fastest
//
// so that the GE x KC planes one CTA needs for a tracer are one contiguous 8 KB block
// and consecutive tracers of the same (group, level chunk) follow each other.
// Internal element order follows the space-filling curve, so a group is a compact
// patch and most DSS neighbours of an element sit in the same group.
#pragma once
#include <cstddef>
#include <cstdint>

namespace tse {

constexpr int NP = 4;
constexpr int NPSQ = 16;
constexpr int NLEV = 72;
constexpr int KC = 4;            // levels per chunk
constexpr int NKC = NLEV / KC;   // 18
#ifndef TSE_GE
#define TSE_GE 16
#endif
constexpr int GE = TSE_GE;       // elements per group (16 = a 4x4 patch at ne = 2^k; 4 measured slower: 2x the halo, 4x the CTA prologues)
constexpr int GPL = GE * KC;     // planes per (group, chunk, tracer) = 64

// direction order of control_mod.F90:173-181 (0-based)
enum { WEST = 0, EAST = 1, SOUTH = 2, NORTH = 3, SWEST = 4, SEAST = 5, NWEST = 6, NEAST = 7 };

// plane index of (element e, tracer q, level k) in a tracer field with Q tracers
__host__ __device__ inline size_t qplane(int e, int q, int k, int Q) {
  const int g = e / GE, el = e % GE, kc = k / KC, kk = k % KC;
  return ((((size_t)g * NKC + kc) * Q + q) * GE + el) * KC + kk;
}
// plane index of (element e, level k) in a level field
__host__ __device__ inline size_t lplane(int e, int k) {
  const int g = e / GE, el = e % GE, kc = k / KC, kk = k % KC;
  return (((size_t)g * NKC + kc) * GE + el) * KC + kk;
}
// plane index of (element e, level k, component c) in vn0
__host__ __device__ inline size_t vplane(int e, int k, int c) {
  const int g = e / GE, el = e % GE, kc = k / KC, kk = k % KC;
  return ((((size_t)g * NKC + kc) * 2 + c) * GE + el) * KC + kk;
}

// DSS gather table: 20 slots per element in the reference's unpack order
// (edge_mod.F90:678-735): S0..3, E0..3, N0..3, W0..3, SW, SE, NE, NW.
// entry >= 0 : (source element << 4) | source node, element in internal order
// entry == -1: no neighbour (missing corner element at a cube corner)
// entry <= -2: ghost slot -(entry+2) of the halo receive buffer (neighbour on another GPU)
constexpr int NSLOT = 20;

}  // namespace tse
