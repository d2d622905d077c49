// ORACLE (test infrastructure only -- never linked into the product path).
//
// CPU restatement, in plain C++17/fp64, of the finite-element side of the
// ExaConstit hot path for linear (p=1) hexahedra with the 2x2x2 Gauss rule.
// Every routine cites the reference file:line whose arithmetic it follows
// (paths relative to /root/reference).  Array layouts are the reference's:
//   E-vector        X(a,i,e)       -> x[e*24 + i*8 + a]
//   jacobian        J(i,s,q,e)     -> jac[((e*8+q)*3 + s)*3 + i]
//   shape gradients G(a,s,q)       -> G[q*24 + s*8 + a]
//   quadrature fns  f(c,q,e)       -> qf[(e*8+q)*vdim + c]
//   matGrad (6x6)   K(i,j,q,e)     -> k[(e*8+q)*36 + j*6 + i]
//   matGradPA (81)  C(i,j,k,l,q,e) -> c[(e*8+q)*81 + ((l*3+k)*3+j)*3 + i]
//   pa_dmat         D(e,q,i,k,l,n) -> row-major
//   ea_data         E(r,c,e)       -> ea[e*576 + c*24 + r]
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

namespace orc {

constexpr int NN = 8;   // nodes per hex
constexpr int NQ = 8;   // quadrature points per hex
constexpr int ED = 24;  // element dofs

// MFEM hexahedron vertex (= H1 p=1 NATIVE dof) reference coordinates.
static const int kHexVert[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0},
                                   {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};

// Shape-function gradients of the trilinear hex at the 2x2x2 Gauss points of
// the unit cube (x fastest) and the weights (1/8 each).  Restates
// el.CalcDShape(ip, DSh) filled column-major per point at
// src/mechanics_operator.cpp:249-260 and src/mechanics_integrators.cpp:184-197.
inline void hex8_dshape(double* G /*192*/, double* W /*8*/) {
  const double g[2] = {0.5 - 0.5 / std::sqrt(3.0), 0.5 + 0.5 / std::sqrt(3.0)};
  for (int qz = 0; qz < 2; ++qz)
    for (int qy = 0; qy < 2; ++qy)
      for (int qx = 0; qx < 2; ++qx) {
        const int q = qx + 2 * qy + 4 * qz;
        const double xi[3] = {g[qx], g[qy], g[qz]};
        W[q] = 0.125;
        for (int a = 0; a < 8; ++a) {
          double n[3], d[3];
          for (int s = 0; s < 3; ++s) {
            n[s] = kHexVert[a][s] ? xi[s] : 1.0 - xi[s];
            d[s] = kHexVert[a][s] ? 1.0 : -1.0;
          }
          G[q * 24 + 0 * 8 + a] = d[0] * n[1] * n[2];
          G[q * 24 + 1 * 8 + a] = n[0] * d[1] * n[2];
          G[q * 24 + 2 * 8 + a] = n[0] * n[1] * d[2];
        }
      }
}

// Cartesian voxel mesh as Mesh::MakeCartesian3D(..., sfc_ordering=false)
// builds it (src/mechanics_driver.cpp:247-253): elements and vertices
// lexicographic, x fastest.  e2n is in NATIVE (hex vertex) order; coords are
// stored byNODES (xxx..yyy..zzz) like the MFEM nodal grid function
// (src/system_driver.cpp:346-349).
inline void voxel_mesh(int nx, int ny, int nz, double lx, double ly, double lz,
                       int* e2n, double* coords) {
  const int px = nx + 1, py = ny + 1, pz = nz + 1;
  const long nn = (long)px * py * pz;
  for (int k = 0; k < pz; ++k)
    for (int j = 0; j < py; ++j)
      for (int i = 0; i < px; ++i) {
        const long n = ((long)k * py + j) * px + i;
        coords[0 * nn + n] = lx * i / nx;
        coords[1 * nn + n] = ly * j / ny;
        coords[2 * nn + n] = lz * k / nz;
      }
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        const long e = ((long)k * ny + j) * nx + i;
        for (int a = 0; a < 8; ++a)
          e2n[e * 8 + a] = (int)((((long)k + kHexVert[a][2]) * py + (j + kHexVert[a][1])) * px +
                                 (i + kHexVert[a][0]));
      }
}

// L-vector (byNODES) -> E-vector, ElementRestriction::Mult with NATIVE
// ordering (src/mechanics_operator.cpp:345, src/mechanics_operator_ext.cpp:150).
inline void gather(long ne, long nn, const int* e2n, const double* xL, double* xE) {
#pragma omp parallel for schedule(static)
  for (long e = 0; e < ne; ++e)
    for (int i = 0; i < 3; ++i)
      for (int a = 0; a < 8; ++a) xE[e * 24 + i * 8 + a] = xL[i * nn + e2n[e * 8 + a]];
}

// E-vector -> L-vector scatter-add, ElementRestriction::MultTranspose
// (src/mechanics_operator_ext.cpp:156).  yL must be zeroed by the caller.
inline void scatter_add(long ne, long nn, const int* e2n, const double* yE, double* yL) {
#pragma omp parallel for schedule(static)
  for (long e = 0; e < ne; ++e)
    for (int i = 0; i < 3; ++i)
      for (int a = 0; a < 8; ++a) {
#pragma omp atomic
        yL[i * nn + e2n[e * 8 + a]] += yE[e * 24 + i * 8 + a];
      }
}

// Jacobians at the quadrature points from nodal coordinates in E-vector form.
// Stands in for mesh->GetGeometricFactors(JACOBIANS) + the transposing copy of
// NonlinearMechOperator::SetupJacobianTerms (src/mechanics_operator.cpp:350-391):
// J(i,s,q,e) = sum_a X(a,i,e) G(a,s,q).
inline void jacobians(long ne, const double* G, const double* xE, double* jac) {
#pragma omp parallel for schedule(static)
  for (long e = 0; e < ne; ++e)
    for (int q = 0; q < 8; ++q)
      for (int s = 0; s < 3; ++s)
        for (int i = 0; i < 3; ++i) {
          double v = 0.0;
          for (int a = 0; a < 8; ++a) v += xE[e * 24 + i * 8 + a] * G[q * 24 + s * 8 + a];
          jac[((e * 8 + q) * 3 + s) * 3 + i] = v;
        }
}

// adj[3*r+c] = adj(J)(r,c) and det(J), written exactly as
// src/mechanics_integrators.cpp:252-270,444-457.
inline double adjugate(const double* Jq, double* adj) {
  const double J11 = Jq[0], J21 = Jq[1], J31 = Jq[2];
  const double J12 = Jq[3], J22 = Jq[4], J32 = Jq[5];
  const double J13 = Jq[6], J23 = Jq[7], J33 = Jq[8];
  const double detJ = J11 * (J22 * J33 - J32 * J23) - J21 * (J12 * J33 - J32 * J13) +
                      J31 * (J12 * J23 - J22 * J13);
  adj[0] = (J22 * J33) - (J23 * J32);
  adj[1] = (J32 * J13) - (J12 * J33);
  adj[2] = (J12 * J23) - (J22 * J13);
  adj[3] = (J31 * J23) - (J21 * J33);
  adj[4] = (J11 * J33) - (J13 * J31);
  adj[5] = (J21 * J13) - (J11 * J23);
  adj[6] = (J21 * J32) - (J31 * J22);
  adj[7] = (J31 * J12) - (J11 * J32);
  adj[8] = (J11 * J22) - (J12 * J21);
  return detJ;
}

// exaconstit::kernel::grad_calc (src/mechanics_kernels.cpp:7-78).
// grad(i,t,q,e) -> out[(e*8+q)*9 + t*3 + i]; accumulates (+=) like the reference.
inline void grad_calc(long ne, const double* jac, const double* G, const double* xE, double* out) {
#pragma omp parallel for schedule(static)
  for (long e = 0; e < ne; ++e)
    for (int q = 0; q < 8; ++q) {
      double adj[9];
      const double detJ = adjugate(&jac[(e * 8 + q) * 9], adj);
      const double c = 1.0 / detJ;
      for (int t = 0; t < 3; ++t)
        for (int s = 0; s < 3; ++s)
          for (int r = 0; r < 8; ++r)
            for (int i = 0; i < 3; ++i)
              out[(e * 8 + q) * 9 + t * 3 + i] +=
                  xE[e * 24 + i * 8 + r] * G[q * 24 + s * 8 + r] * (c * adj[3 * s + t]);
    }
}

static const int kVoigt[3][3] = {{0, 5, 4}, {5, 1, 3}, {4, 3, 2}};

// ExaModel::TransformMatGradTo4D (src/mechanics_model.cpp:949-1061).
inline void transform_matgrad_4d(long npts, const double* k36, double* c81) {
#pragma omp parallel for schedule(static)
  for (long p = 0; p < npts; ++p)
    for (int l = 0; l < 3; ++l)
      for (int k = 0; k < 3; ++k)
        for (int j = 0; j < 3; ++j)
          for (int i = 0; i < 3; ++i)
            c81[p * 81 + ((l * 3 + k) * 3 + j) * 3 + i] =
                k36[p * 36 + kVoigt[k][l] * 6 + kVoigt[i][j]];
}

// ExaNLFIntegrator::AssemblePA (src/mechanics_integrators.cpp:240-312):
// Dres(j,k,q,e) = W_q sum_m adj(J)(j,m) sigma(m,k) -> d[((e*8+q)*3 + k)*3 + j].
inline void assemble_pa(long ne, const double* jac, const double* W, const double* stress, double* d) {
#pragma omp parallel for schedule(static)
  for (long e = 0; e < ne; ++e)
    for (int q = 0; q < 8; ++q) {
      double adj[9];
      adjugate(&jac[(e * 8 + q) * 9], adj);
      const double* S = &stress[(e * 8 + q) * 6];
      for (int k = 0; k < 3; ++k)
        for (int j = 0; j < 3; ++j) {
          double v = 0.0;
          // order of the three products as in the reference: m = k-row of sigma
          v = S[kVoigt[0][k]] * adj[3 * j + 0] + S[kVoigt[1][k]] * adj[3 * j + 1] +
              S[kVoigt[2][k]] * adj[3 * j + 2];
          d[((e * 8 + q) * 3 + k) * 3 + j] = v * W[q];
        }
    }
}

// ExaNLFIntegrator::AddMultPA (src/mechanics_integrators.cpp:545-555).
inline void addmult_pa(long ne, const double* G, const double* d, double* yE) {
#pragma omp parallel for schedule(static)
  for (long e = 0; e < ne; ++e)
    for (int q = 0; q < 8; ++q)
      for (int k = 0; k < 3; ++k)
        for (int j = 0; j < 3; ++j)
          for (int i = 0; i < 8; ++i)
            yE[e * 24 + k * 8 + i] += G[q * 24 + j * 8 + i] * d[((e * 8 + q) * 3 + k) * 3 + j];
}

// ExaNLFIntegrator::AssembleGradPA (src/mechanics_integrators.cpp:425-511).
// A(r,c) = adj(J)(c,r); D(e,q,i,k,l,n) = (dt W/detJ) sum_{j,m} A(j,i) C(j,k,l,m) A(m,n).
inline void assemble_grad_pa(long ne, double dt, const double* jac, const double* W,
                             const double* c81, double* D) {
#pragma omp parallel for schedule(static)
  for (long e = 0; e < ne; ++e)
    for (int q = 0; q < 8; ++q) {
      double adj[9];
      const double detJ = adjugate(&jac[(e * 8 + q) * 9], adj);
      const double c_detJ = 1.0 / detJ * W[q] * dt;
      auto A = [&](int r, int c) { return adj[r + 3 * c]; };
      const double* C = &c81[(e * 8 + q) * 81];
      auto C4 = [&](int i, int j, int k, int l) { return C[((l * 3 + k) * 3 + j) * 3 + i]; };
      double* Dq = &D[(e * 8 + q) * 81];
      for (int x = 0; x < 81; ++x) Dq[x] = 0.0;
      for (int n = 0; n < 3; ++n)
        for (int m = 0; m < 3; ++m)
          for (int l = 0; l < 3; ++l)
            for (int i = 0; i < 3; ++i)
              for (int k = 0; k < 3; ++k)
                Dq[((i * 3 + k) * 3 + l) * 3 + n] +=
                    (A(0, i) * C4(0, k, l, m) + A(1, i) * C4(1, k, l, m) + A(2, i) * C4(2, k, l, m)) *
                    A(m, n);
      for (int x = 0; x < 81; ++x) Dq[x] *= c_detJ;
    }
}

// ExaNLFIntegrator::AddMultGradPA (src/mechanics_integrators.cpp:592-620).
inline void addmult_grad_pa(long ne, const double* G, const double* D, const double* xE, double* yE) {
#pragma omp parallel for schedule(static)
  for (long e = 0; e < ne; ++e)
    for (int q = 0; q < 8; ++q) {
      const double* Dq = &D[(e * 8 + q) * 81];
      double T[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
          for (int k = 0; k < 8; ++k) {
            const double gx = G[q * 24 + j * 8 + k] * xE[e * 24 + i * 8 + k];
            for (int b = 0; b < 3; ++b)
              for (int a = 0; a < 3; ++a) T[a + 3 * b] += Dq[((a * 3 + b) * 3 + i) * 3 + j] * gx;
          }
      for (int k = 0; k < 3; ++k)
        for (int j = 0; j < 3; ++j)
          for (int i = 0; i < 8; ++i) yE[e * 24 + k * 8 + i] += G[q * 24 + j * 8 + i] * T[j + 3 * k];
    }
}

// Physical shape gradients scaled by det J: b[c] = sum_s G(a,s,q) adj(J)(s,c)
// (the column-major view A(r,c)=adj(J)(c,r) at src/mechanics_integrators.cpp:673,703-713).
inline void bvec(const double* G, const double* adj, int q, int a, double* b) {
  for (int c = 0; c < 3; ++c)
    b[c] = G[q * 24 + 0 * 8 + a] * adj[0 + c] + G[q * 24 + 1 * 8 + a] * adj[3 + c] +
           G[q * 24 + 2 * 8 + a] * adj[6 + c];
}

// ExaNLFIntegrator::AssembleGradDiagonalPA (src/mechanics_integrators.cpp:668-746).
inline void assemble_grad_diag_pa(long ne, double dt, const double* jac, const double* W,
                                  const double* G, const double* k36, double* dE) {
#pragma omp parallel for schedule(static)
  for (long e = 0; e < ne; ++e)
    for (int q = 0; q < 8; ++q) {
      double adj[9];
      const double detJ = adjugate(&jac[(e * 8 + q) * 9], adj);
      const double c = 1.0 / detJ * W[q] * dt;
      const double* Kq = &k36[(e * 8 + q) * 36];
      auto K = [&](int i, int j) { return Kq[j * 6 + i]; };
      for (int a = 0; a < 8; ++a) {
        double b[3];
        bvec(G, adj, q, a, b);
        for (int I = 0; I < 3; ++I) {
          double v = 0.0;
          for (int p = 0; p < 3; ++p) {
            double w = 0.0;
            for (int r = 0; r < 3; ++r) w += b[r] * K(kVoigt[I][p], kVoigt[I][r]);
            v += b[p] * w;
          }
          dE[e * 24 + I * 8 + a] += c * v;
        }
      }
    }
}

// ExaNLFIntegrator::AssembleEA (src/mechanics_integrators.cpp:849-1015):
// E(l+8I, k+8Kc, e) += c * sum_{p,r} g_p(l) K(v(I,p), v(Kc,r)) b_r(k).
inline void assemble_ea(long ne, double dt, const double* jac, const double* W, const double* G,
                        const double* k36, double* ea) {
#pragma omp parallel for schedule(static)
  for (long e = 0; e < ne; ++e)
    for (int q = 0; q < 8; ++q) {
      double adj[9];
      const double detJ = adjugate(&jac[(e * 8 + q) * 9], adj);
      const double c = 1.0 / detJ * W[q] * dt;
      const double* Kq = &k36[(e * 8 + q) * 36];
      auto K = [&](int i, int j) { return Kq[j * 6 + i]; };
      for (int k = 0; k < 8; ++k) {
        double b[3];
        bvec(G, adj, q, k, b);
        // kk[I][Kc][p] = c * sum_r K(v(I,p), v(Kc,r)) b_r
        double kk[3][3][3];
        for (int I = 0; I < 3; ++I)
          for (int Kc = 0; Kc < 3; ++Kc)
            for (int p = 0; p < 3; ++p)
              kk[I][Kc][p] = c * (b[0] * K(kVoigt[I][p], kVoigt[Kc][0]) +
                                  b[1] * K(kVoigt[I][p], kVoigt[Kc][1]) +
                                  b[2] * K(kVoigt[I][p], kVoigt[Kc][2]));
        for (int l = 0; l < 8; ++l) {
          double g[3];
          bvec(G, adj, q, l, g);
          for (int I = 0; I < 3; ++I)
            for (int Kc = 0; Kc < 3; ++Kc)
              ea[e * 576 + (k + 8 * Kc) * 24 + (l + 8 * I)] +=
                  g[0] * kk[I][Kc][0] + g[1] * kk[I][Kc][1] + g[2] * kk[I][Kc][2];
        }
      }
    }
}

// EANonlinearMechOperatorGradExt::TMult inner loop
// (src/mechanics_operator_ext.cpp:303-314): Y(j,e) += sum_i A(i,j,e) X(i,e).
inline void ea_mult(long ne, const double* ea, const double* xE, double* yE) {
#pragma omp parallel for schedule(static)
  for (long e = 0; e < ne; ++e)
    for (int j = 0; j < 24; ++j) {
      double res = 0.0;
      for (int i = 0; i < 24; ++i) res += ea[e * 576 + j * 24 + i] * xE[e * 24 + i];
      yE[e * 24 + j] += res;
    }
}

// EA AssembleDiagonal element part (src/mechanics_operator_ext.cpp:246-252).
inline void ea_diag(long ne, const double* ea, double* dE) {
#pragma omp parallel for schedule(static)
  for (long e = 0; e < ne; ++e)
    for (int j = 0; j < 24; ++j) dE[e * 24 + j] = ea[e * 576 + j * 24 + j];
}

// ---------------------------------------------------------------- B-bar ----
// ICExaNLFIntegrator::AssemblePA element-average shape gradients
// (src/mechanics_integrators.cpp:1895-1952): eDS(a,c,e) -> eds[e*24 + c*8 + a].
inline void ic_assemble_eds(long ne, const double* jac, const double* W, const double* G, double* eds) {
#pragma omp parallel for schedule(static)
  for (long e = 0; e < ne; ++e) {
    double vol = 0.0;
    for (int x = 0; x < 24; ++x) eds[e * 24 + x] = 0.0;
    for (int q = 0; q < 8; ++q) {
      double adj[9];
      const double detJ = adjugate(&jac[(e * 8 + q) * 9], adj);
      vol += W[q] * detJ;
      for (int a = 0; a < 8; ++a) {
        double b[3];
        bvec(G, adj, q, a, b);
        for (int c = 0; c < 3; ++c) eds[e * 24 + c * 8 + a] += W[q] * b[c];
      }
    }
    const double ivol = 1.0 / vol;
    for (int x = 0; x < 24; ++x) eds[e * 24 + x] *= ivol;
  }
}

// The 6x3 B-bar block of node a at a quadrature point, rows in Voigt order
// (src/mechanics_integrators.cpp:1289-1306, 2050-2083).
inline void bbar_block(const double* G, const double* adj, double idetJ, const double* eds_e, int q,
                       int a, double Bb[6][3]) {
  double b[3];
  bvec(G, adj, q, a, b);
  for (int c = 0; c < 3; ++c) b[c] *= idetJ;
  const double i3 = 1.0 / 3.0;
  const double b4 = i3 * (eds_e[0 * 8 + a] - b[0]), b5 = b4 + b[0];
  const double b6 = i3 * (eds_e[1 * 8 + a] - b[1]), b7 = b6 + b[1];
  const double b8 = i3 * (eds_e[2 * 8 + a] - b[2]), b9 = b8 + b[2];
  const double t[6][3] = {{b5, b6, b8}, {b4, b7, b8}, {b4, b6, b9},
                          {0.0, b[2], b[1]}, {b[2], 0.0, b[0]}, {b[1], b[0], 0.0}};
  std::memcpy(Bb, t, sizeof(t));
}

// ICExaNLFIntegrator::AddMultPA (src/mechanics_integrators.cpp:2011-2085).
inline void ic_addmult_pa(long ne, const double* jac, const double* W, const double* G,
                          const double* eds, const double* stress, double* yE) {
#pragma omp parallel for schedule(static)
  for (long e = 0; e < ne; ++e)
    for (int q = 0; q < 8; ++q) {
      double adj[9];
      const double detJ = adjugate(&jac[(e * 8 + q) * 9], adj);
      const double c = detJ * W[q];
      const double* S = &stress[(e * 8 + q) * 6];
      for (int a = 0; a < 8; ++a) {
        double Bb[6][3];
        bbar_block(G, adj, 1.0 / detJ, &eds[e * 24], q, a, Bb);
        for (int I = 0; I < 3; ++I) {
          double v = 0.0;
          for (int R = 0; R < 6; ++R) v += Bb[R][I] * S[R];
          yE[e * 24 + I * 8 + a] += c * v;
        }
      }
    }
}

// ICExaNLFIntegrator::AssembleEA (src/mechanics_integrators.cpp:1250-1601):
// E(l+8I, k+8Kc, e) += dt W detJ * sum_{R,S} Bbar_l(R,I) K(R,S) Bbar_k(S,Kc).
inline void ic_assemble_ea(long ne, double dt, const double* jac, const double* W, const double* G,
                           const double* eds, const double* k36, double* ea) {
#pragma omp parallel for schedule(static)
  for (long e = 0; e < ne; ++e)
    for (int q = 0; q < 8; ++q) {
      double adj[9];
      const double detJ = adjugate(&jac[(e * 8 + q) * 9], adj);
      const double c = detJ * W[q] * dt;
      const double* Kq = &k36[(e * 8 + q) * 36];
      double Bb[8][6][3];
      for (int a = 0; a < 8; ++a) bbar_block(G, adj, 1.0 / detJ, &eds[e * 24], q, a, Bb[a]);
      for (int k = 0; k < 8; ++k) {
        double KB[6][3];  // c * K(R,:) . Bbar_k(:,Kc)
        for (int R = 0; R < 6; ++R)
          for (int Kc = 0; Kc < 3; ++Kc) {
            double v = 0.0;
            for (int S = 0; S < 6; ++S) v += Kq[S * 6 + R] * Bb[k][S][Kc];
            KB[R][Kc] = c * v;
          }
        for (int l = 0; l < 8; ++l)
          for (int I = 0; I < 3; ++I)
            for (int Kc = 0; Kc < 3; ++Kc) {
              double v = 0.0;
              for (int R = 0; R < 6; ++R) v += Bb[l][R][I] * KB[R][Kc];
              ea[e * 576 + (k + 8 * Kc) * 24 + (l + 8 * I)] += v;
            }
      }
    }
}

// ICExaNLFIntegrator::AssembleGradDiagonalPA (src/mechanics_integrators.cpp:1655-1803):
// the diagonal of the B-bar element matrix.
inline void ic_assemble_grad_diag_pa(long ne, double dt, const double* jac, const double* W,
                                     const double* G, const double* eds, const double* k36,
                                     double* dE) {
#pragma omp parallel for schedule(static)
  for (long e = 0; e < ne; ++e)
    for (int q = 0; q < 8; ++q) {
      double adj[9];
      const double detJ = adjugate(&jac[(e * 8 + q) * 9], adj);
      const double c = detJ * W[q] * dt;
      const double* Kq = &k36[(e * 8 + q) * 36];
      for (int a = 0; a < 8; ++a) {
        double Bb[6][3];
        bbar_block(G, adj, 1.0 / detJ, &eds[e * 24], q, a, Bb);
        for (int I = 0; I < 3; ++I) {
          double v = 0.0;
          for (int R = 0; R < 6; ++R) {
            double w = 0.0;
            for (int S = 0; S < 6; ++S) w += Kq[S * 6 + R] * Bb[S][I];
            v += Bb[R][I] * w;
          }
          dE[e * 24 + I * 8 + a] += c * v;
        }
      }
    }
}

// ComputeVolAvgTensor (src/mechanics_kernels.hpp:51-133): sum_{e,q} detJ W f and the volume.
inline void vol_sum(long ne, int vdim, const double* jac, const double* W, const double* qf,
                    double* sums, double* vol) {
  for (int c = 0; c < vdim; ++c) sums[c] = 0.0;
  double v = 0.0;
  std::vector<double> acc(vdim, 0.0);
#pragma omp parallel
  {
    std::vector<double> loc(vdim, 0.0);
    double lv = 0.0;
#pragma omp for schedule(static) nowait
    for (long e = 0; e < ne; ++e)
      for (int q = 0; q < 8; ++q) {
        double adj[9];
        const double w = adjugate(&jac[(e * 8 + q) * 9], adj) * W[q];
        lv += w;
        for (int c = 0; c < vdim; ++c) loc[c] += w * qf[(e * 8 + q) * vdim + c];
      }
#pragma omp critical
    {
      v += lv;
      for (int c = 0; c < vdim; ++c) acc[c] += loc[c];
    }
  }
  *vol = v;
  for (int c = 0; c < vdim; ++c) sums[c] = acc[c];
}

}  // namespace orc
