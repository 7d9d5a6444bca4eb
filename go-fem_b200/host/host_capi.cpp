// host_capi.cpp -- extern "C" wrappers over the C++ host mirror (gofem.hpp) for the Python harness.
#include "gofem.hpp"
#include "../../include/femb200_host.h"
#include <cstring>
#include <memory>

using namespace gofem;

struct gofem_ga {
  std::unique_ptr<fem::GeneralAssembler> ga;
  Error last;
};

static std::unique_ptr<elements::IsoBase> make_elem(int kind, int quad_order, int node_dofs) {
  std::unique_ptr<elements::IsoBase> e;
  switch (kind) {
  case GOFEM_TETRA4: e.reset(new elements::Tetra4()); break;
  case GOFEM_TETRA10: e.reset(new elements::Tetra10()); break;
  case GOFEM_HEXA8: e.reset(new elements::Hexa8()); break;
  case GOFEM_HEXA20: e.reset(new elements::Hexa20()); break;
  case GOFEM_QUAD4: e.reset(new elements::Quad4()); break;
  case GOFEM_QUAD8: e.reset(new elements::Quad8()); break;
  case GOFEM_TRIANGLE3: e.reset(new elements::Triangle3()); break;
  case GOFEM_TRIANGLE6: e.reset(new elements::Triangle6()); break;
  default: return e;
  }
  e->QuadratureOrder = quad_order;
  e->NodeDofs = (fem::DofsFlag)node_dofs;
  return e;
}

static bool make_constituter(int ckind, double p0, double p1, solids::isoconstituter &c) {
  switch (ckind) {
  case GOFEM_ISO_SOLID3D: c = solids::Isotropic{p0, p1}.Solid3D(); return true;
  case GOFEM_ISO_AXISYM: c = solids::Isotropic{p0, p1}.Axisymmetric(); return true;
  case GOFEM_ISO_PLANESTRESS: c = solids::Isotropic{p0, p1}.PlaneStess(); return true;
  case GOFEM_ISO_PLANESTRAIN: c = solids::Isotropic{p0, p1}.PlaneStrain(); return true;
  case GOFEM_COND_SOLID3D: c = solids::IsotropicConductivity{p0}.Solid3D(); return true;
  case GOFEM_COND_PLANE: c = solids::IsotropicConductivity{p0}.Plane(); return true;
  case GOFEM_COND_AXISYM: c = solids::IsotropicConductivity{p0}.Axisymmetric(); return true;
  }
  return false;
}

extern "C" {

int gofem_elem_info(int kind, int quad_order, int node_dofs, int32_t *len_nodes, int32_t *dofs_flag, double *iso_nodes,
                    int32_t *nqp, double *qpos, double *qw, char *name, int name_cap) {
  auto e = make_elem(kind, quad_order, node_dofs);
  if (!e) return -1;
  if (len_nodes) *len_nodes = e->LenNodes();
  if (dofs_flag) *dofs_flag = e->Dofs();
  if (iso_nodes) {
    auto n = e->IsoparametricNodes();
    for (size_t i = 0; i < n.size(); ++i) { iso_nodes[3 * i] = n[i].X; iso_nodes[3 * i + 1] = n[i].Y; iso_nodes[3 * i + 2] = n[i].Z; }
  }
  std::vector<r3::Vec> p; std::vector<double> w;
  e->Quadrature(p, w);
  if (nqp) *nqp = (int32_t)p.size();
  for (size_t i = 0; i < p.size() && i < 216; ++i) {
    if (qpos) { qpos[3 * i] = p[i].X; qpos[3 * i + 1] = p[i].Y; qpos[3 * i + 2] = p[i].Z; }
    if (qw) qw[i] = w[i];
  }
  if (name && name_cap > 0) { std::string s = e->String(); std::strncpy(name, s.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  return 0;
}

int gofem_elem_basis(int kind, const double v[3], double *N, double *dN, int32_t *len_dN) {
  auto e = make_elem(kind, 0, 0);
  if (!e) return -1;
  r3::Vec p{v[0], v[1], v[2]};
  auto n = e->Basis(p), d = e->BasisDiff(p);
  if (N) std::memcpy(N, n.data(), sizeof(double) * n.size());
  if (dN) std::memcpy(dN, d.data(), sizeof(double) * d.size());
  if (len_dN) *len_dN = (int32_t)d.size();
  return 0;
}

int gofem_constitutive(int ckind, double p0, double p1, double *C, int32_t *dimC, int32_t *bkind) {
  solids::isoconstituter c;
  if (!make_constituter(ckind, p0, p1, c)) return -1;
  std::vector<double> m; int r = 0, cc = 0;
  Error err = c.Constitutive(m, r, cc);
  if (err) return 1;
  if (C) std::memcpy(C, m.data(), sizeof(double) * m.size());
  if (dimC) *dimC = r;
  if (bkind) *bkind = c.BKind();
  return 0;
}

int gofem_new_general_assembler(int device, const double *nodes_xyz, int64_t n_nodes, int model_dofs_flag, gofem_ga **out) {
  if (!out) return -1;
  std::vector<r3::Vec> nodes((size_t)n_nodes);
  for (int64_t i = 0; i < n_nodes; ++i) nodes[(size_t)i] = r3::Vec{nodes_xyz[3 * i], nodes_xyz[3 * i + 1], nodes_xyz[3 * i + 2]};
  gofem_ga *g = new gofem_ga();
  g->ga.reset(new fem::GeneralAssembler(nodes, (fem::DofsFlag)model_dofs_flag, device));
  *out = g;
  if (!g->ga->InitError().empty()) { g->last = Error::New(g->ga->InitError()); return FEMB200_ERR_CUDA; }
  return 0;
}

void gofem_free(gofem_ga *ga) { delete ga; }

int gofem_add_isoparametric(gofem_ga *g, int elem_kind, int quad_order, int node_dofs, int ckind, double p0, double p1,
                            const int64_t *conn, const int32_t *elem_lens, int64_t n_elem, const double orient_x[3]) {
  auto e = make_elem(elem_kind, quad_order, node_dofs);
  solids::isoconstituter c;
  if (!e || !make_constituter(ckind, p0, p1, c)) { g->last = Error::New("bad element/constituter kind"); return -1; }
  const int nn = e->LenNodes();
  std::vector<int64_t> off;
  if (elem_lens) { off.resize((size_t)n_elem + 1, 0); for (int64_t i = 0; i < n_elem; ++i) off[i + 1] = off[i] + elem_lens[i]; }
  r3::Vec ox{};
  if (orient_x) ox = r3::Vec{orient_x[0], orient_x[1], orient_x[2]};
  g->last = g->ga->AddIsoparametric(*e, c, n_elem, [&](int64_t i, std::vector<int64_t> &elem, r3::Vec &x, r3::Vec &y) {
    if (elem_lens) elem.assign(conn + off[i], conn + off[i + 1]);
    else elem.assign(conn + i * nn, conn + (i + 1) * nn);
    x = ox; y = r3::Vec{};
  });
  return g->last ? 1 : 0;
}

int gofem_add_isoparametric_flat(gofem_ga *g, int elem_kind, int quad_order, int node_dofs, int ckind, double p0, double p1,
                                 const int64_t *conn, int64_t n_elem) {
  auto e = make_elem(elem_kind, quad_order, node_dofs);
  solids::isoconstituter c;
  if (!e || !make_constituter(ckind, p0, p1, c)) { g->last = Error::New("bad element/constituter kind"); return -1; }
  g->last = g->ga->AddIsoparametricFlat(*e, c, n_elem, conn);
  return g->last ? 1 : 0;
}

int gofem_isoparametric_strains(gofem_ga *g, const double *u, int64_t n_u, int elem_kind, int quad_order, int node_dofs,
                                int ckind, double p0, double p1, const int64_t *conn, int64_t n_elem, double *strains) {
  auto e = make_elem(elem_kind, quad_order, node_dofs);
  solids::isoconstituter c;
  if (!e || !make_constituter(ckind, p0, p1, c)) { g->last = Error::New("bad element/constituter kind"); return -1; }
  const int nn = e->LenNodes();
  std::vector<double> uu(u, u + n_u);
  g->last = g->ga->IsoparametricStrains(uu, *e, c, n_elem,
      [&](int64_t i, std::vector<int64_t> &elem, r3::Vec &, r3::Vec &) { elem.assign(conn + i * nn, conn + (i + 1) * nn); },
      [&](int64_t iele, const double *s, int n) { std::memcpy(strains + iele * n, s, sizeof(double) * n); });
  return g->last ? 1 : 0;
}

const char *gofem_last_error(const gofem_ga *g, int32_t *is_panic) {
  if (is_panic) *is_panic = g->last.is_panic ? 1 : 0;
  return g->last.msg.c_str();
}
int64_t gofem_total_dofs(const gofem_ga *g) { return g->ga->TotalDofs(); }
int64_t gofem_ksolid_nnz(gofem_ga *g) { return g->ga->Ksolid()->NonZero(); }
double gofem_ksolid_at(gofem_ga *g, int64_t i, int64_t j) { return g->ga->Ksolid()->At(i, j); }
int gofem_ksolid_csr(gofem_ga *g, int64_t *indptr, int32_t *indices, double *values) {
  lap::Sparse *K = g->ga->Ksolid();
  if (indptr) std::memcpy(indptr, K->Indptr().data(), sizeof(int64_t) * K->Indptr().size());
  if (indices) std::memcpy(indices, K->Indices().data(), sizeof(int32_t) * K->Indices().size());
  if (values) std::memcpy(values, K->Values().data(), sizeof(double) * K->Values().size());
  return 0;
}
int gofem_ksolid_mulvec(gofem_ga *g, const double *x, double *y) {
  std::vector<double> xx(x, x + g->ga->TotalDofs()), yy;
  g->ga->Ksolid()->MulVec(xx, yy);
  std::memcpy(y, yy.data(), sizeof(double) * yy.size());
  return 0;
}
femb200_assembler *gofem_device_assembler(gofem_ga *g) { return g->ga->Device(); }
femb200_ctx *gofem_device_ctx(gofem_ga *g) { return g->ga->Context(); }

int64_t gofem_fixity_dofs(int model_dofs_flag, int64_t n_nodes, const uint16_t *fixed_flags, int get_fixed, int64_t *out) {
  fem::Fixity f((fem::DofsFlag)model_dofs_flag, n_nodes);
  for (int64_t i = 0; i < n_nodes; ++i) f.Fix(i, fixed_flags[i]);
  auto v = get_fixed ? f.FixedDofs() : f.FreeDofs();
  if (out) std::memcpy(out, v.data(), sizeof(int64_t) * v.size());
  return (int64_t)v.size();
}

} // extern "C"
