/* Minimal C host for the C-ABI: reads packed tables written by tools/export_tables.py, solves 64 modes,
 * prints P(k).  Build:  gcc examples/evolve_c_abi.c -Iinclude -Ldisco-eb_b200/discoeb_b200 -ldiscoeb_b200 -lm -o evolve
 *         run :  LD_LIBRARY_PATH=disco-eb_b200/discoeb_b200 ./evolve tables.bin                                   */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "discoeb_b200.h"

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s tables.bin  (int32 nth, int32 nnu, scalars[24], tables[...])\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("open"); return 2; }
  int32_t nth, nnu;
  if (fread(&nth, 4, 1, f) != 1 || fread(&nnu, 4, 1, f) != 1) return 2;
  deb_dims d = {0};
  d.ncosmo = 1; d.nk = 64; d.nout = 1; d.lmaxg = d.lmaxgp = d.lmaxr = 11; d.lmaxnu = 8; d.nqmax = 3;
  d.nth = nth; d.nnu = nnu; d.max_steps = 2048; d.power_idx = 4;
  deb_ctrl c = {1e-4, 1e-4, 0.25, 0.80, 0.0, 20.0, 0.3, 0.9};
  size_t tl = deb_table_len(&d);
  double scalars[DEB_NSCAL];
  double* tables = (double*)malloc(tl * sizeof(double));
  if (fread(scalars, 8, DEB_NSCAL, f) != DEB_NSCAL || fread(tables, 8, tl, f) != tl) return 2;
  fclose(f);
  double k[64], aout[1] = {1.0}, y[64 * 20], pk[64], tau_out[1];
  int32_t status[64], nsteps[64], nacc[64];
  for (int i = 0; i < 64; ++i) k[i] = 1e-4 * pow(1e5, i / 63.0);
  float ms = 0;
  int rc = deb_evolve_host_f64(&d, &c, scalars, tables, k, aout, y, pk, tau_out, status, nsteps, nacc, 0, &ms);
  if (rc) { fprintf(stderr, "deb_evolve_host_f64: %s\n", deb_strerror(rc)); return 1; }
  printf("# kernel %.2f ms, tau(a=1) = %.3f Mpc\n# k [1/Mpc]   P_m(k) [Mpc^3]   steps\n", ms, tau_out[0]);
  for (int i = 0; i < 64; i += 4) printf("%.5e  %.5e  %d%s\n", k[i], pk[i], nsteps[i], status[i] ? "  (failed)" : "");
  free(tables);
  return 0;
}
