/* C11 client of include/tse.h: builds a small cubed-sphere case with the host mesh library's C entry points, runs the
 * device-side DCMIP driver for one remap cycle through the C ABI -- no ctypes, no C++ -- and prints the order-independent
 * diagnostics.  tests/test_c_abi.py compares the output with the same run through the Python binding: equal field fingerprints
 * prove that the struct layouts of tse.h are what the library reads.
 *
 *   abi_check <ne> <qsize> <test 11|12> <tstep> <nu_q> <vcoord.txt>      vcoord.txt: 73 hyai, 73 hybi, 72 hyam, 72 hybm */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tse.h"

/* host mesh library (transport_se_b200/csrc/tse_mesh.cpp, extern "C" section) */
void tse_gll(double* pts4, double* wts4, double* dvv16, double* mp16);
void* tse_mesh_create(int ne);
void tse_mesh_destroy(void* m);
int tse_mesh_nelem(void* m);
void tse_mesh_get(void* m, double* lat, double* lon, double* D, double* Dinv, double* metdet, double* rmetdet, double* spheremp,
                  double* rspheremp, int* nbr, int* nbr_dir, int* rev, int* sfc, int* face_ie_je);
void tse_mesh_sfc_partition(void* m, int nparts, int* owner);
void* tse_local_view_create(void* m, const int* owner, int rank, int nranks);
void tse_local_view_destroy(void* v);
void tse_local_view_sizes(void* v, int* nelemd, int* nbuf, int* ncycles);
void tse_local_view_get(void* v, int* gid, int* putmap, int* getmap, int* reverse, int* cyc_rank, int* cyc_ptr, int* cyc_len);

#define CHECK(call)                                                        \
  do {                                                                     \
    if ((call) != 0) {                                                     \
      fprintf(stderr, "abi_check: %s: %s\n", #call, tse_last_error());     \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 7) {
    fprintf(stderr, "usage: abi_check ne qsize test tstep nu_q vcoord.txt\n");
    return 2;
  }
  const int ne = atoi(argv[1]), qsize = atoi(argv[2]), test = atoi(argv[3]);
  const double tstep = atof(argv[4]), nu_q = atof(argv[5]);
  double hy[73 + 73 + 72 + 72];
  FILE* f = fopen(argv[6], "r");
  if (!f) return 2;
  for (int i = 0; i < 290; ++i)
    if (fscanf(f, "%lf", &hy[i]) != 1) return 2;
  fclose(f);

  void* mesh = tse_mesh_create(ne);
  const int nelem = tse_mesh_nelem(mesh);
  const size_t n16 = (size_t)nelem * 16;
  double* buf = malloc(sizeof(double) * n16 * (6 + 8));
  double *lat = buf, *lon = lat + n16, *metdet = lon + n16, *rmetdet = metdet + n16, *spheremp = rmetdet + n16, *rspheremp = spheremp + n16,
         *D = rspheremp + n16, *Dinv = D + 4 * n16;
  int* nbr = malloc(sizeof(int) * (size_t)nelem * 8);
  int* sfc = malloc(sizeof(int) * (size_t)nelem);
  tse_mesh_get(mesh, lat, lon, D, Dinv, metdet, rmetdet, spheremp, rspheremp, nbr, NULL, NULL, sfc, NULL);
  int* owner = calloc((size_t)nelem, sizeof(int));
  void* view = tse_local_view_create(mesh, owner, 0, 1);
  int nelemd = 0, nbuf = 0, ncycles = 0;
  tse_local_view_sizes(view, &nelemd, &nbuf, &ncycles);
  int* gid = malloc(sizeof(int) * (size_t)nelemd);
  int* maps = malloc(sizeof(int) * (size_t)nelemd * 24);
  int *putmap = maps, *getmap = maps + (size_t)nelemd * 8, *reverse = getmap + (size_t)nelemd * 8;
  int cyc[3] = {0, 0, 0};
  tse_local_view_get(view, gid, putmap, getmap, reverse, cyc, cyc + 1, cyc + 2);

  double pts[4], wts[4], dvv[16], mp[16];
  tse_gll(pts, wts, dvv, mp);

  tse_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.ne = ne; cfg.nelemd = nelemd; cfg.qsize = qsize; cfg.qsize_d = qsize; cfg.nlev = 72; cfg.np = 4;
  cfg.rsplit = 3; cfg.qsplit = 1; cfg.limiter_option = 8; cfg.hypervis_order = 2; cfg.hypervis_subcycle_q = 1;
  cfg.vert_remap_q_alg = 0; cfg.nu_q = nu_q; cfg.device = -1;
  const tse_geometry geom = {spheremp, rspheremp, metdet, rmetdet, Dinv, lat, lon};
  const tse_connectivity conn = {putmap, getmap, reverse, nbuf, sfc, 0, NULL, NULL, NULL};
  const tse_hvcoord hv = {hy, hy + 73, hy + 146, hy + 218, 100000.0};
  tse_handle h = NULL;
  CHECK(tse_init(&cfg, &geom, &conn, &hv, dvv, &h));
  CHECK(tse_dcmip_init(h, test));
  int nstep = 0;
  CHECK(tse_prim_run_subcycle(h, tstep, &nstep));
  const int tl = nstep % 2 == 0 ? 1 : 2;
  double* mass = malloc(sizeof(double) * (size_t)qsize * 3);
  unsigned long long* hash = malloc(sizeof(unsigned long long) * (size_t)qsize);
  CHECK(tse_diag_mass(h, tl, mass));
  CHECK(tse_diag_qminmax(h, tl, mass + qsize, mass + 2 * qsize));
  CHECK(tse_diag_field_hash(h, tl, hash));
  printf("nstep %d\n", nstep);
  for (int q = 0; q < qsize; ++q) printf("tracer %d mass %a qmin %a qmax %a hash %016llx\n", q, mass[q], mass[qsize + q], mass[2 * qsize + q], hash[q]);
  CHECK(tse_finalize(h));
  tse_local_view_destroy(view);
  tse_mesh_destroy(mesh);
  free(buf); free(nbr); free(sfc); free(owner); free(gid); free(maps); free(mass); free(hash);
  return 0;
}
