/*
 * b200_vtk_binary.h -- binary twin of the reference's particle VTK writer.
 *
 * particle_results_vtk__InOutFun__ (InOutFun/Outputs/WriteVtk.c:95-268) prints every value with
 * fprintf("%.20g"): ~25 bytes and a libc formatting call per double; at 10^6..10^7 particles one results step costs
 * seconds to minutes of host time and tens of GB.  This writer emits the SAME legacy-VTK datasets (same names, same
 * order, same 2D -> 3D padding rules, read from the same Fields) in the BINARY flavour of the format (big-endian raw
 * values), 8 bytes per double.  It is host code against the reference's own headers and is called by the B200 scheme
 * shims in place of the ASCII writer when NLPS_B200_VTK_BINARY=1; it runs on the copy that nlps_b200_download_end
 * delivered, while the GPU steps on.
 *
 * Datasets covered: POINTS / CELLS / CELL_TYPES, X_GC, MASS, DENSITY, ELEM_i, MatIdx, VELOCITY, ACCELERATION,
 * DISPLACEMENT, STRESS, P, DEFORMATION-GRADIENT, Energy-Potential, Energy-Kinetic, EPS.  When the deck asks for any
 * other dataset (X_EC, damage, strain, Green-Lagrange, metric, plastic deformation gradient / jacobian, Pw) the function
 * returns 1 without touching the disk and the caller falls back to the reference's writer: the output is never
 * silently incomplete.
 */
#ifndef B200_VTK_BINARY_H
#define B200_VTK_BINARY_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t b200_be64(double v) {
  uint64_t u;
  memcpy(&u, &v, 8);
  return __builtin_bswap64(u);
}
static inline uint32_t b200_be32(int v) { return __builtin_bswap32((uint32_t)v); }

/* n rows of `cols` doubles padded to `pad` columns with zeros, big-endian */
static void b200_vtk_rows(FILE *f, const double *a, int n, int cols, int stride, int pad, uint64_t *buf) {
  const uint64_t zero = b200_be64(0.0);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < pad; j++) buf[(size_t)i * pad + j] = (j < cols) ? b200_be64(a[(size_t)i * stride + j]) : zero;
  }
  fwrite(buf, 8, (size_t)n * pad, f);
}
/* 3 x 3 tensors from d x d blocks (row-major in a stride-wide row); slot33 >= 0: column of the out-of-plane entry */
static void b200_vtk_tensors(FILE *f, const double *a, int n, int d, int stride, int slot33, uint64_t *buf) {
  const uint64_t zero = b200_be64(0.0);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < 3; j++)
      for (int k = 0; k < 3; k++) {
        uint64_t v = zero;
        if (j < d && k < d) v = b200_be64(a[(size_t)i * stride + j * d + k]);
        else if (j == 2 && k == 2 && slot33 >= 0) v = b200_be64(a[(size_t)i * stride + slot33]);
        buf[(size_t)i * 9 + j * 3 + k] = v;
      }
  fwrite(buf, 8, (size_t)n * 9, f);
}
static void b200_vtk_ints(FILE *f, const int *a, int n, uint32_t *buf) {
  for (int i = 0; i < n; i++) buf[i] = b200_be32(a[i]);
  fwrite(buf, 4, (size_t)n, f);
}

/* returns 0 written, 1 not written (dataset outside the covered set, or the file cannot be opened) */
static int b200_particle_results_vtk_binary(Particle MPM_Mesh, int TimeStep_i, int ResultsTimeStep_) {
  if (Out_element_coordinates || Out_damage || Out_eigenvalues_stress || Out_water_pressure || Out_Pw || Out_dPw_dt ||
      Out_strain || Out_eigenvalues_strain || Out_green_lagrange || Out_plastic_deformation_gradient || Out_Metric ||
      Out_plastic_jacobian || Out_Von_Mises)
    return 1;
  const int d = NumberDimensions, T = (NumberDimensions == 2) ? 5 : 9, n = MPM_Mesh.NumGP;
  char name[10000];
  sprintf(name, "%s/%s_%i.vtk", OutputDir, OutputParticlesFile, TimeStep_i);
  FILE *f = fopen(name, "wb");
  if (!f) return 1;
  uint64_t *buf = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1) * 9);
  uint32_t *ibuf = (uint32_t *)buf;
  const Fields *P = &MPM_Mesh.Phi;
  fprintf(f, "# vtk DataFile Version 3.0\nResults time step %i\nBINARY\nDATASET UNSTRUCTURED_GRID\n", ResultsTimeStep_);
  fprintf(f, "POINTS %i double\n", n);
  b200_vtk_rows(f, P->x_GC.nV, n, d, d, 3, buf);
  fprintf(f, "\nCELLS %i %i\n", n, 2 * n);
  for (int i = 0; i < n; i++) { ibuf[2 * i] = b200_be32(1); ibuf[2 * i + 1] = b200_be32(i); }
  fwrite(ibuf, 4, (size_t)2 * n, f);
  fprintf(f, "\nCELL_TYPES %i\n", n);
  for (int i = 0; i < n; i++) ibuf[i] = b200_be32(1);
  fwrite(ibuf, 4, (size_t)n, f);
  fprintf(f, "\nPOINT_DATA %i\n", n);
  if (Out_global_coordinates) { fprintf(f, "VECTORS X_GC double\n"); b200_vtk_rows(f, P->x_GC.nV, n, d, d, 3, buf); fprintf(f, "\n"); }
  if (Out_mass) { fprintf(f, "SCALARS MASS double\nLOOKUP_TABLE default\n"); b200_vtk_rows(f, P->mass.nV, n, 1, 1, 1, buf); fprintf(f, "\n"); }
  if (Out_density) { fprintf(f, "SCALARS DENSITY double\nLOOKUP_TABLE default\n"); b200_vtk_rows(f, P->rho.nV, n, 1, 1, 1, buf); fprintf(f, "\n"); }
  if (Out_nodal_idx) { fprintf(f, "SCALARS ELEM_i int\nLOOKUP_TABLE default\n"); b200_vtk_ints(f, MPM_Mesh.I0, n, ibuf); fprintf(f, "\n"); }
  if (Out_material_idx) { fprintf(f, "SCALARS MatIdx int\nLOOKUP_TABLE default\n"); b200_vtk_ints(f, MPM_Mesh.MatIdx, n, ibuf); fprintf(f, "\n"); }
  if (Out_velocity) { fprintf(f, "VECTORS VELOCITY double\n"); b200_vtk_rows(f, P->vel.nV, n, d, d, 3, buf); fprintf(f, "\n"); }
  if (Out_acceleration) { fprintf(f, "VECTORS ACCELERATION double\n"); b200_vtk_rows(f, P->acc.nV, n, d, d, 3, buf); fprintf(f, "\n"); }
  if (Out_displacement) { fprintf(f, "VECTORS DISPLACEMENT double\n"); b200_vtk_rows(f, P->dis.nV, n, d, d, 3, buf); fprintf(f, "\n"); }
  if (Out_stress) { /* WriteVtk.c:568-588: the 33 entry of a 2D stress is slot 4 */
    fprintf(f, "TENSORS STRESS double\n");
    b200_vtk_tensors(f, P->Stress.nV, n, d, T, d == 2 ? 4 : -1, buf);
    fprintf(f, "\n");
  }
  if (Out_volumetric_stress) { /* :591-607 */
    fprintf(f, "SCALARS P double\nLOOKUP_TABLE default\n");
    for (int i = 0; i < n; i++) {
      const double *s = &P->Stress.nV[(size_t)i * T];
      const double pr = (d == 2) ? (1.0 / 3.0) * (s[0] + s[3] + s[4]) : (1.0 / 3.0) * (s[0] + s[4] + s[8]);
      buf[i] = b200_be64(pr);
    }
    fwrite(buf, 8, (size_t)n, f);
    fprintf(f, "\n");
  }
  if (Out_deformation_gradient) { /* :651-668: no out-of-plane entry */
    fprintf(f, "TENSORS DEFORMATION-GRADIENT double\n");
    b200_vtk_tensors(f, P->F_n.nV, n, d, T, -1, buf);
    fprintf(f, "\n");
  }
  if (Out_energy) { /* :816-841 */
    fprintf(f, "SCALARS Energy-Potential double\nLOOKUP_TABLE default\n");
    b200_vtk_rows(f, P->W, n, 1, 1, 1, buf);
    fprintf(f, "\nSCALARS Energy-Kinetic double\nLOOKUP_TABLE default\n");
    for (int i = 0; i < n; i++) {
      double k = 0.0;
      for (int j = 0; j < d; j++) k += P->vel.nV[(size_t)i * d + j] * P->vel.nV[(size_t)i * d + j];
      buf[i] = b200_be64(0.5 * k * P->mass.nV[i]);
    }
    fwrite(buf, 8, (size_t)n, f);
    fprintf(f, "\n");
  }
  if (Out_EPS) { fprintf(f, "SCALARS EPS double\nLOOKUP_TABLE default\n"); b200_vtk_rows(f, P->EPS_n, n, 1, 1, 1, buf); fprintf(f, "\n"); }
  free(buf);
  return fclose(f) == 0 ? 0 : 1;
}

/* what the scheme shims call at a results step */
static void b200_write_particle_results(Particle MPM_Mesh, int TimeStep_i, int ResultsTimeStep_) {
  const char *s = getenv("NLPS_B200_VTK_BINARY");
  if (s && atoi(s) != 0 && b200_particle_results_vtk_binary(MPM_Mesh, TimeStep_i, ResultsTimeStep_) == 0) return;
  particle_results_vtk__InOutFun__(MPM_Mesh, TimeStep_i, ResultsTimeStep_);
}
#endif /* B200_VTK_BINARY_H */
