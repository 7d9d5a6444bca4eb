// gofem.hpp -- C++ host-side mirror of the go-fem API for the AddIsoparametric path.
//
// The Go toolchain is not available in the build image, so the host layer that a Go user would see
// (packages fem, elements, constitution/solids, plus the few lap.Sparse methods the reference calls)
// is mirrored here in C++ with the same names, argument meaning and error behaviour; the cgo shim a
// maintainer would add is shown in INTEGRATION.md and go-fem_b200/gofem/.  Everything numeric happens
// behind the C ABI of include/femb200.h (CUDA, sm_100a); this layer only evaluates the once-per-call
// tables (assembler_isoparametric.go:29-71), flattens the getElement callback (fem.go:141) and formats
// the reference's error strings.  There is no CPU assembly path.
//
// Shape functions are written from the standard closed forms (tensor-product / serendipity /
// barycentric), keyed by the reference's node numbering -- not transliterated from
// elements/<type>.go; the oracle holds the literal restatement and tests compare the two.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <memory>
#include <string>
#include <vector>
#include "../../include/femb200.h"

namespace gofem {

namespace r3 {
struct Vec { double X = 0, Y = 0, Z = 0; }; // gonum spatial/r3.Vec
inline bool operator==(const Vec &a, const Vec &b) { return a.X == b.X && a.Y == b.Y && a.Z == b.Z; }
inline bool operator!=(const Vec &a, const Vec &b) { return !(a == b); }
} // namespace r3

// Go-style error value: empty message == nil.
struct Error {
  std::string msg;
  bool is_panic = false; // the reference panics (programmer error) instead of returning an error
  explicit operator bool() const { return !msg.empty(); }
  static Error Nil() { return Error{}; }
  static Error New(const std::string &m) { return Error{m, false}; }
  static Error Panic(const std::string &m) { return Error{m, true}; }
};

namespace fem {

// fem.go:68-91
using DofsFlag = uint16_t;
constexpr DofsFlag DofPosX = 1 << 0, DofPosY = 1 << 1, DofPosZ = 1 << 2, DofRotX = 1 << 3, DofRotY = 1 << 4, DofRotZ = 1 << 5;
constexpr DofsFlag DofPos = DofPosX | DofPosY | DofPosZ, DofRot = DofRotX | DofRotY | DofRotZ, Dof6 = DofPos | DofRot;
inline int Count(DofsFlag d) { int c = 0; for (int i = 0; i < 16; ++i) c += (d >> i) & 1; return c; } // fem.go:94-96
inline bool Has(DofsFlag d, DofsFlag q) { return (d & q) == q; }                                       // fem.go:100-102
inline std::string String(DofsFlag d) {                                                                // fem.go:105-118
  if (d == 0) return "none";
  std::string s;
  for (int i = 0; d != 0; ++i) { s += (d & 1) ? std::to_string(i + 1) : std::string("-"); d >>= 1; }
  return s;
}

// fem.go:13-46
struct Element {
  virtual ~Element() = default;
  virtual int LenNodes() const = 0;
  virtual DofsFlag Dofs() const = 0;
};
struct Isoparametric : Element {
  virtual std::vector<r3::Vec> IsoparametricNodes() const = 0;
  virtual void Quadrature(std::vector<r3::Vec> &positions, std::vector<double> &weights) const = 0;
  virtual std::vector<double> Basis(r3::Vec v) const = 0;
  virtual std::vector<double> BasisDiff(r3::Vec v) const = 0; // [d][nn] row major
  virtual std::string String() const = 0;
};

// fem.go:55-65.  BKind() is the tag that replaces the unexported strain func of solids.isoconstituter
// (solids.go:8-12): it names which device-side B filler to run; -1 = user defined (cannot run on the GPU).
struct Constituter {
  virtual ~Constituter() = default;
  virtual Error Constitutive(std::vector<double> &C, int &rows, int &cols) const = 0;
};
struct IsoConstituter : Constituter {
  virtual int BKind() const { return -1; }
};

// fem.go:155-172 (same quirk: the loop runs to popcount(model), unfilled entries stay 0)
inline Error mapdofs(DofsFlag modelDofs, DofsFlag elemDofs, std::vector<int> &dofMapping) {
  const int numModelDofs = Count(modelDofs);
  if (!Has(modelDofs, elemDofs))
    return Error::New("model dofs " + String(modelDofs) + " does not contain all element dofs " + String(elemDofs));
  dofMapping.assign(Count(elemDofs), 0);
  int idm = 0;
  for (int i = 0; i < numModelDofs; ++i)
    if (Has(modelDofs, DofsFlag(1 << i)) && Has(elemDofs, DofsFlag(1 << i))) dofMapping[idm++] = i;
  if (idm == 0 || dofMapping.empty()) return Error::New("element has empty dof mapping");
  return Error::Nil();
}

} // namespace fem

namespace lap {
// The slice of soypat/lap's Sparse that the reference's call sites use: Dims, At, MulVec,
// NonZero, backed by the canonical CSR fetched from the device on first access after a change.
class Sparse {
public:
  void Dims(int64_t &r, int64_t &c) const { r = c = n_; }
  int64_t NonZero() { sync(); return (int64_t)values_.size(); }
  double At(int64_t i, int64_t j) {
    sync();
    if (i < 0 || j < 0 || i >= n_ || j >= n_) return 0.0;
    int64_t lo = indptr_[i], hi = indptr_[i + 1];
    while (lo < hi) { int64_t mid = (lo + hi) / 2; if (indices_[mid] < j) lo = mid + 1; else hi = mid; }
    return (lo < indptr_[i + 1] && indices_[lo] == j) ? values_[lo] : 0.0;
  }
  void MulVec(const std::vector<double> &x, std::vector<double> &y) { // on the device
    y.assign((size_t)n_, 0.0);
    if (ga_) femb200_csr_matvec(ga_, x.data(), y.data());
  }
  const std::vector<int64_t> &Indptr() { sync(); return indptr_; }
  const std::vector<int32_t> &Indices() { sync(); return indices_; }
  const std::vector<double> &Values() { sync(); return values_; }
  // internal
  void bind(femb200_assembler *ga, int64_t n) { ga_ = ga; n_ = n; dirty_ = true; }
  void invalidate() { dirty_ = true; }
private:
  void sync() {
    if (!dirty_) return;
    indptr_.assign((size_t)n_ + 1, 0);
    int64_t nnz = ga_ ? femb200_nnz(ga_) : 0;
    indices_.assign((size_t)nnz, 0);
    values_.assign((size_t)nnz, 0.0);
    if (ga_) femb200_csr_copy(ga_, indptr_.data(), indices_.data(), values_.data());
    dirty_ = false;
  }
  femb200_assembler *ga_ = nullptr;
  int64_t n_ = 0;
  bool dirty_ = true;
  std::vector<int64_t> indptr_;
  std::vector<int32_t> indices_;
  std::vector<double> values_;
};
} // namespace lap

namespace elements {

// 1-D Gauss-Legendre rules, quadrature.go:50-111.  n=4 reproduces the reference's a==b abscissa
// (sqrt((3-sqrt(6/5))/7) used for both) and n=6 its 15-digit constants.
inline bool gaussQuad1D(int n, std::vector<double> &x, std::vector<double> &w) {
  switch (n) {
  case 1: x = {0.0}; w = {2.0}; return true;
  case 2: { const double a = 1.0 / std::sqrt(3.0L); x = {-a, a}; w = {1.0, 1.0}; return true; }
  case 3: { const double a = (double)std::sqrt(0.6L); x = {-a, 0.0, a}; w = {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0}; return true; }
  case 4: {
    const double a = (double)std::sqrt((3.0L - std::sqrt(1.2L)) / 7.0L), b = a;
    const double wa = (double)((18.0L + std::sqrt(30.0L)) / 36.0L), wb = (double)((18.0L - std::sqrt(30.0L)) / 36.0L);
    x = {-b, -a, a, b}; w = {wb, wa, wa, wb}; return true;
  }
  case 5: {
    const double a = (double)(std::sqrt(5.0L - 2.0L * std::sqrt(10.0L / 7.0L)) / 3.0L);
    const double b = (double)(std::sqrt(5.0L + 2.0L * std::sqrt(10.0L / 7.0L)) / 3.0L);
    const double wa = (double)((322.0L + 13.0L * std::sqrt(70.0L)) / 900.0L), wb = (double)((322.0L - 13.0L * std::sqrt(70.0L)) / 900.0L);
    x = {-b, -a, 0.0, a, b}; w = {wb, wa, 128.0 / 225.0, wa, wb}; return true;
  }
  case 6:
    x = {-0.932469514203152, -0.661209386466265, -0.238619186083197, 0.238619186083197, 0.661209386466265, 0.932469514203152};
    w = {0.171324492379170, 0.360761573048139, 0.467913934572691, 0.467913934572691, 0.360761573048139, 0.171324492379170};
    return true;
  }
  return false;
}
// quadrature.go:22-48 (x outermost, z innermost) and :9-18 (2-D: z rule of one point, weights halved)
inline bool uniformGaussQuad(int nx, int ny, int nz, std::vector<r3::Vec> &pos, std::vector<double> &w) {
  std::vector<double> x, wx, y, wy, z, wz;
  if (!gaussQuad1D(nx, x, wx) || !gaussQuad1D(ny, y, wy) || !gaussQuad1D(nz, z, wz)) return false;
  pos.clear(); w.clear();
  for (int i = 0; i < nx; ++i) for (int j = 0; j < ny; ++j) for (int k = 0; k < nz; ++k) {
    w.push_back(wx[i] * wy[j] * wz[k]);
    pos.push_back(r3::Vec{x[i], y[j], z[k]});
  }
  return true;
}
inline bool uniformGaussQuad2d(int nx, int ny, std::vector<r3::Vec> &pos, std::vector<double> &w) {
  if (!uniformGaussQuad(nx, ny, 1, pos, w)) return false;
  for (double &v : w) v /= 2;
  return true;
}
// triangles.go:158-212: tabulated rules (data), weights halved
inline bool triangleQuads(int order, std::vector<r3::Vec> &pos, std::vector<double> &w) {
  switch (order) {
  case 1: pos = {{1.0 / 3, 1.0 / 3, 0}}; w = {1.0}; break;
  case 2: pos = {{1.0 / 6, 2.0 / 3, 0}, {1.0 / 6, 1.0 / 6, 0}, {2.0 / 3, 1.0 / 6, 0}}; w = {1.0 / 3, 1.0 / 3, 1.0 / 3}; break;
  case 3: pos = {{1.0 / 3, 1.0 / 3, 0}, {0.2, 0.6, 0}, {0.2, 0.2, 0}, {0.6, 0.2, 0}};
          w = {-0.5625, 0.520833333333333, 0.520833333333333, 0.520833333333333}; break;
  case 4: pos = {{0.445948490915965, 0.108103018168070, 0}, {0.445948490915965, 0.445948490915965, 0}, {0.108103018168070, 0.445948490915965, 0},
                 {0.091576213509771, 0.816847572980459, 0}, {0.091576213509771, 0.091576213509771, 0}, {0.816847572980459, 0.091576213509771, 0}};
          w = {0.223381589678011, 0.223381589678011, 0.223381589678011, 0.109951743655322, 0.109951743655322, 0.109951743655322}; break;
  default: return false;
  }
  for (double &v : w) v *= 0.5;
  return true;
}

// Base for the reference's element structs: exported fields QuadratureOrder, NodeDofs.
struct IsoBase : fem::Isoparametric {
  int QuadratureOrder = 0;
  fem::DofsFlag NodeDofs = 0;
  bool quad_ok = true; // set false when the reference would panic in Quadrature()
};

// ---- Tetra4 (tetra4.go:10-55): L = (1-x-y-z, x, y, z); one point, weight 1 (8-Q1); ignores NodeDofs
struct Tetra4 : IsoBase {
  int LenNodes() const override { return 4; }
  fem::DofsFlag Dofs() const override { return fem::DofPos; }
  std::vector<r3::Vec> IsoparametricNodes() const override { return {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}}; }
  std::vector<double> Basis(r3::Vec v) const override { return {1 - v.X - v.Y - v.Z, v.X, v.Y, v.Z}; }
  std::vector<double> BasisDiff(r3::Vec) const override { return {-1, 1, 0, 0, -1, 0, 1, 0, -1, 0, 0, 1}; }
  void Quadrature(std::vector<r3::Vec> &p, std::vector<double> &w) const override { p = {{.25, .25, .25}}; w = {1.0}; }
  std::string String() const override { return "TETRA4"; }
};

// ---- Tetra10 (tetra10.go): corners L_i(2L_i-1); mid-edges 4 L_i L_j on edges (0-1)(0-2)(0-3)(1-2)(2-3)(1-3)
struct Tetra10 : IsoBase {
  int LenNodes() const override { return 10; }
  fem::DofsFlag Dofs() const override { return NodeDofs ? NodeDofs : fem::DofPos; }
  std::vector<r3::Vec> IsoparametricNodes() const override {
    return {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {.5, 0, 0}, {0, .5, 0}, {0, 0, .5}, {.5, .5, 0}, {0, .5, .5}, {.5, 0, .5}};
  }
  std::vector<double> Basis(r3::Vec v) const override { // tetra10.go:45-58: the reference's expansion, term by term (same rounding)
    std::vector<double> N(10);
    const double X = v.X, Y = v.Y, Z = v.Z;
    N[0] = 2*X*X + 4*X*Y + 4*X*Z - 3*X + 2*Y*Y + 4*Y*Z - 3*Y + 2*Z*Z - 3*Z + 1;
    N[1] = 2*X*X - X;
    N[2] = 2*Y*Y - Y;
    N[3] = 2*Z*Z - Z;
    N[4] = 4*X - 4*X*Y - 4*X*Z - 4*X*X;
    N[5] = 4*Y - 4*X*Y - 4*Y*Z - 4*Y*Y;
    N[6] = 4*Z - 4*X*Z - 4*Y*Z - 4*Z*Z;
    N[7] = 4 * X * Y;
    N[8] = 4 * Y * Z;
    N[9] = 4 * X * Z;
    return N;
  }
  std::vector<double> BasisDiff(r3::Vec v) const override { // tetra10.go:60-96: the reference's expansion, term by term (same rounding)
    std::vector<double> d(30);
    const double X = v.X, Y = v.Y, Z = v.Z;
    d[0] = 4*X + 4*Y + 4*Z - 3;
    d[1] = 4*X - 1;
    d[2] = 0; d[3] = 0;
    d[4] = 4 - 4*Y - 4*Z - 8*X;
    d[5] = -4 * Y;
    d[6] = -4 * Z;
    d[7] = 4 * Y;
    d[8] = 0;
    d[9] = 4 * Z;
    d[10] = 4*X + 4*Y + 4*Z - 3;
    d[11] = 0;
    d[12] = 4*Y - 1;
    d[13] = 0;
    d[14] = -4 * X;
    d[15] = 4 - 8*Y - 4*Z - 4*X;
    d[16] = -4 * Z;
    d[17] = 4 * X;
    d[18] = 4 * Z;
    d[19] = 0;
    d[20] = 4*X + 4*Y + 4*Z - 3;
    d[21] = 0; d[22] = 0;
    d[23] = 4*Z - 1;
    d[24] = -4 * X;
    d[25] = -4 * Y;
    d[26] = 4 - 4*Y - 8*Z - 4*X;
    d[27] = 0;
    d[28] = 4 * Y;
    d[29] = 4 * X;
    return d;
  }
  void Quadrature(std::vector<r3::Vec> &p, std::vector<double> &w) const override { // tetra10.go:98-112, weights 1/4 (8-Q1)
    const double a = (double)((5.0L + 3.0L * std::sqrt(5.0L)) / 20.0L), b = (double)((5.0L - std::sqrt(5.0L)) / 20.0L);
    p = {{a, b, b}, {b, b, b}, {b, b, a}, {b, a, b}};
    w = {0.25, 0.25, 0.25, 0.25};
  }
  std::string String() const override { return "TETRA10"; }
};

// ---- tensor-product / serendipity families keyed by the reference's node tables
struct Hexa8 : IsoBase {
  static constexpr double P[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}}; // hexa8.go:32-43
  int LenNodes() const override { return 8; }
  fem::DofsFlag Dofs() const override { return NodeDofs ? NodeDofs : fem::DofPos; }
  std::vector<r3::Vec> IsoparametricNodes() const override { std::vector<r3::Vec> n; for (auto &p : P) n.push_back({p[0], p[1], p[2]}); return n; }
  std::vector<double> Basis(r3::Vec v) const override { // hexa8.go:46-57: the reference's expansion, term by term (same rounding)
    std::vector<double> N(8);
    const double X = v.X, Y = v.Y, Z = v.Z;
    N[0] = (Z*Y - Y - X - Z + Z*X + Y*X - Z*Y*X + 1) / 8;
    N[1] = (X - Y - Z + Z*Y - Z*X - Y*X + Z*Y*X + 1) / 8;
    N[2] = (Y - Z + X - Z*Y - Z*X + Y*X - Z*Y*X + 1) / 8;
    N[3] = (Y - Z - X - Z*Y + Z*X - Y*X + Z*Y*X + 1) / 8;
    N[4] = (Z - Y - X - Z*Y - Z*X + Y*X + Z*Y*X + 1) / 8;
    N[5] = (Z - Y + X - Z*Y + Z*X - Y*X - Z*Y*X + 1) / 8;
    N[6] = (Z + Y + X + Z*Y + Z*X + Y*X + Z*Y*X + 1) / 8;
    N[7] = (Z + Y - X + Z*Y - Z*X - Y*X - Z*Y*X + 1) / 8;
    return N;
  }
  std::vector<double> BasisDiff(r3::Vec v) const override { // hexa8.go:67-96: the reference's expansion, term by term (same rounding)
    std::vector<double> d(24);
    const double X = v.X, Y = v.Y, Z = v.Z;
    d[0] = (Z + Y - (Z * Y) - 1) / 8;
    d[1] = ((Z * Y) - Y - Z + 1) / 8;
    d[2] = (Y - Z - (Z * Y) + 1) / 8;
    d[3] = (Z - Y + (Z * Y) - 1) / 8;
    d[4] = (Y - Z + (Z * Y) - 1) / 8;
    d[5] = (Z - Y - (Z * Y) + 1) / 8;
    d[6] = (Z + Y + (Z * Y) + 1) / 8;
    d[7] = (-Z - Y - (Z * Y) - 1) / 8;
    d[8] = (Z + X - (Z * X) - 1) / 8;
    d[9] = (Z - X + (Z * X) - 1) / 8;
    d[10] = (X - Z - (Z * X) + 1) / 8;
    d[11] = ((Z * X) - X - Z + 1) / 8;
    d[12] = (X - Z + (Z * X) - 1) / 8;
    d[13] = (-Z - X - (Z * X) - 1) / 8;
    d[14] = (Z + X + (Z * X) + 1) / 8;
    d[15] = (Z - X - (Z * X) + 1) / 8;
    d[16] = (Y + X - (Y * X) - 1) / 8;
    d[17] = (Y - X + (Y * X) - 1) / 8;
    d[18] = (-Y - X - (Y * X) - 1) / 8;
    d[19] = (X - Y + (Y * X) - 1) / 8;
    d[20] = ((Y * X) - X - Y + 1) / 8;
    d[21] = (X - Y - (Y * X) + 1) / 8;
    d[22] = (Y + X + (Y * X) + 1) / 8;
    d[23] = (Y - X - (Y * X) + 1) / 8;
    return d;
  }
  void Quadrature(std::vector<r3::Vec> &p, std::vector<double> &w) const override { // hexa8.go:98-108
    int q = QuadratureOrder == 0 ? 2 : QuadratureOrder;
    if (!uniformGaussQuad(q, q, q, p, w)) { p.clear(); w.clear(); }
  }
  std::string String() const override { return "HEXA8"; }
};

struct Hexa20 : IsoBase {
  static constexpr double P[20][3] = { // hexa20.go:33-56
      {-1, -1, -1}, {-1, 1, -1}, {1, 1, -1}, {1, -1, -1}, {-1, -1, 1}, {-1, 1, 1}, {1, 1, 1}, {1, -1, 1},
      {-1, 0, -1}, {0, 1, -1}, {1, 0, -1}, {0, -1, -1}, {-1, 0, 1}, {0, 1, 1}, {1, 0, 1}, {0, -1, 1},
      {-1, -1, 0}, {-1, 1, 0}, {1, 1, 0}, {1, -1, 0}};
  int LenNodes() const override { return 20; }
  fem::DofsFlag Dofs() const override { return NodeDofs ? NodeDofs : fem::DofPos; }
  std::vector<r3::Vec> IsoparametricNodes() const override { std::vector<r3::Vec> n; for (auto &p : P) n.push_back({p[0], p[1], p[2]}); return n; }
  std::vector<double> Basis(r3::Vec v) const override { // hexa20.go:59-86: the reference's expansion, term by term (same rounding)
    std::vector<double> N(20);
    const double X = v.X, Y = v.Y, Z = v.Z;
    double x2 = X * X, y2 = Y * Y, z2 = Z * Z;
    N[0] = ((Y - 1) * (X - 1) * (Z - 1) * (Y + X + Z + 2)) / 8;
    N[1] = -((Y + 1) * (X - 1) * (Z - 1) * (X - Y + Z + 2)) / 8;
    N[2] = -((Y + 1) * (X + 1) * (Z - 1) * (Y + X - Z - 2)) / 8;
    N[3] = -((Y - 1) * (X + 1) * (Z - 1) * (Y - X + Z + 2)) / 8;
    N[4] = -((Y - 1) * (X - 1) * (Z + 1) * (Y + X - Z + 2)) / 8;
    N[5] = -((Y + 1) * (X - 1) * (Z + 1) * (Y - X + Z - 2)) / 8;
    N[6] = ((Y + 1) * (X + 1) * (Z + 1) * (Y + X + Z - 2)) / 8;
    N[7] = ((Y - 1) * (X + 1) * (Z + 1) * (Y - X - Z + 2)) / 8;
    N[8] = -((y2 - 1) * (X - 1) * (Z - 1)) / 4;
    N[9] = ((x2 - 1) * (Y + 1) * (Z - 1)) / 4;
    N[10] = ((y2 - 1) * (X + 1) * (Z - 1)) / 4;
    N[11] = -((x2 - 1) * (Y - 1) * (Z - 1)) / 4;
    N[12] = ((y2 - 1) * (X - 1) * (Z + 1)) / 4;
    N[13] = -((x2 - 1) * (Y + 1) * (Z + 1)) / 4;
    N[14] = -((y2 - 1) * (X + 1) * (Z + 1)) / 4;
    N[15] = ((x2 - 1) * (Y - 1) * (Z + 1)) / 4;
    N[16] = -((z2 - 1) * (Y - 1) * (X - 1)) / 4;
    N[17] = ((z2 - 1) * (Y + 1) * (X - 1)) / 4;
    N[18] = -((z2 - 1) * (Y + 1) * (X + 1)) / 4;
    N[19] = ((z2 - 1) * (Y - 1) * (X + 1)) / 4;
    return N;
  }
  std::vector<double> BasisDiff(r3::Vec v) const override { // hexa20.go:96-163: the reference's expansion, term by term (same rounding)
    std::vector<double> d(60);
    const double X = v.X, Y = v.Y, Z = v.Z;
    double x2 = X * X, y2 = Y * Y, z2 = Z * Z;
    d[0] = ((Y - 1) * (Z - 1) * (Y + 2*X + Z + 1)) / 8;
    d[1] = -((Y + 1) * (Z - 1) * (2*X - Y + Z + 1)) / 8;
    d[2] = -((Y + 1) * (Z - 1) * (Y + 2*X - Z - 1)) / 8;
    d[3] = -((Y - 1) * (Z - 1) * (Y - 2*X + Z + 1)) / 8;
    d[4] = -((Y - 1) * (Z + 1) * (Y + 2*X - Z + 1)) / 8;
    d[5] = -((Y + 1) * (Z + 1) * (Y - 2*X + Z - 1)) / 8;
    d[6] = ((Y + 1) * (Z + 1) * (Y + 2*X + Z - 1)) / 8;
    d[7] = ((Y - 1) * (Z + 1) * (Y - 2*X - Z + 1)) / 8;
    d[8] = -((y2 - 1) * (Z - 1)) / 4;
    d[9] = (X * (Y + 1) * (Z - 1)) / 2;
    d[10] = ((y2 - 1) * (Z - 1)) / 4;
    d[11] = -(X * (Y - 1) * (Z - 1)) / 2;
    d[12] = ((y2 - 1) * (Z + 1)) / 4;
    d[13] = -(X * (Y + 1) * (Z + 1)) / 2;
    d[14] = -((y2 - 1) * (Z + 1)) / 4;
    d[15] = (X * (Y - 1) * (Z + 1)) / 2;
    d[16] = -((z2 - 1) * (Y - 1)) / 4;
    d[17] = ((z2 - 1) * (Y + 1)) / 4;
    d[18] = -((z2 - 1) * (Y + 1)) / 4;
    d[19] = ((z2 - 1) * (Y - 1)) / 4;
    d[20] = ((X - 1) * (Z - 1) * (2*Y + X + Z + 1)) / 8;
    d[21] = -((X - 1) * (Z - 1) * (X - 2*Y + Z + 1)) / 8;
    d[22] = -((X + 1) * (Z - 1) * (2*Y + X - Z - 1)) / 8;
    d[23] = -((X + 1) * (Z - 1) * (2*Y - X + Z + 1)) / 8;
    d[24] = -((X - 1) * (Z + 1) * (2*Y + X - Z + 1)) / 8;
    d[25] = -((X - 1) * (Z + 1) * (2*Y - X + Z - 1)) / 8;
    d[26] = ((X + 1) * (Z + 1) * (2*Y + X + Z - 1)) / 8;
    d[27] = ((X + 1) * (Z + 1) * (2*Y - X - Z + 1)) / 8;
    d[28] = -(Y * (X - 1) * (Z - 1)) / 2;
    d[29] = ((x2 - 1) * (Z - 1)) / 4;
    d[30] = (Y * (X + 1) * (Z - 1)) / 2;
    d[31] = -((x2 - 1) * (Z - 1)) / 4;
    d[32] = (Y * (X - 1) * (Z + 1)) / 2;
    d[33] = -((x2 - 1) * (Z + 1)) / 4;
    d[34] = -(Y * (X + 1) * (Z + 1)) / 2;
    d[35] = ((x2 - 1) * (Z + 1)) / 4;
    d[36] = -((z2 - 1) * (X - 1)) / 4;
    d[37] = ((z2 - 1) * (X - 1)) / 4;
    d[38] = -((z2 - 1) * (X + 1)) / 4;
    d[39] = ((z2 - 1) * (X + 1)) / 4;
    d[40] = ((Y - 1) * (X - 1) * (Y + X + 2*Z + 1)) / 8;
    d[41] = -((Y + 1) * (X - 1) * (X - Y + 2*Z + 1)) / 8;
    d[42] = -((Y + 1) * (X + 1) * (Y + X - 2*Z - 1)) / 8;
    d[43] = -((Y - 1) * (X + 1) * (Y - X + 2*Z + 1)) / 8;
    d[44] = -((Y - 1) * (X - 1) * (Y + X - 2*Z + 1)) / 8;
    d[45] = -((Y + 1) * (X - 1) * (Y - X + 2*Z - 1)) / 8;
    d[46] = ((Y + 1) * (X + 1) * (Y + X + 2*Z - 1)) / 8;
    d[47] = ((Y - 1) * (X + 1) * (Y - X - 2*Z + 1)) / 8;
    d[48] = -((y2 - 1) * (X - 1)) / 4;
    d[49] = ((x2 - 1) * (Y + 1)) / 4;
    d[50] = ((y2 - 1) * (X + 1)) / 4;
    d[51] = -((x2 - 1) * (Y - 1)) / 4;
    d[52] = ((y2 - 1) * (X - 1)) / 4;
    d[53] = -((x2 - 1) * (Y + 1)) / 4;
    d[54] = -((y2 - 1) * (X + 1)) / 4;
    d[55] = ((x2 - 1) * (Y - 1)) / 4;
    d[56] = -(Z * (Y - 1) * (X - 1)) / 2;
    d[57] = (Z * (Y + 1) * (X - 1)) / 2;
    d[58] = -(Z * (Y + 1) * (X + 1)) / 2;
    d[59] = (Z * (Y - 1) * (X + 1)) / 2;
    return d;
  }
  void Quadrature(std::vector<r3::Vec> &p, std::vector<double> &w) const override { // hexa20.go:166-176
    int q = QuadratureOrder == 0 ? 3 : QuadratureOrder;
    if (!uniformGaussQuad(q, q, q, p, w)) { p.clear(); w.clear(); }
  }
  std::string String() const override { return "HEXA20"; }
};

struct Quad4 : IsoBase {
  static constexpr double P[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}}; // quad4.go:35-42
  int LenNodes() const override { return 4; }
  fem::DofsFlag Dofs() const override { return NodeDofs ? NodeDofs : fem::DofsFlag(fem::DofPosX | fem::DofPosY); }
  std::vector<r3::Vec> IsoparametricNodes() const override { std::vector<r3::Vec> n; for (auto &p : P) n.push_back({p[0], p[1], 0}); return n; }
  std::vector<double> Basis(r3::Vec v) const override {
    std::vector<double> N(4);
    for (int a = 0; a < 4; ++a) N[a] = (1 + v.X * P[a][0]) * (1 + v.Y * P[a][1]) / 4;
    return N;
  }
  std::vector<double> BasisDiff(r3::Vec v) const override {
    std::vector<double> d(8);
    for (int a = 0; a < 4; ++a) { d[a] = P[a][0] * (1 + v.Y * P[a][1]) / 4; d[4 + a] = (1 + v.X * P[a][0]) * P[a][1] / 4; }
    return d;
  }
  int order() const { return QuadratureOrder <= 0 ? 2 : QuadratureOrder; } // quad4.go:82-88
  void Quadrature(std::vector<r3::Vec> &p, std::vector<double> &w) const override { if (!uniformGaussQuad2d(order(), order(), p, w)) { p.clear(); w.clear(); } }
  std::string String() const override { return "Quad43d(order=" + std::to_string(order()) + ")"; }
};

struct Quad8 : IsoBase {
  static constexpr double P[8][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}, {0, -1}, {1, 0}, {0, 1}, {-1, 0}}; // quad8.go:36-47
  int LenNodes() const override { return 8; }
  fem::DofsFlag Dofs() const override { return NodeDofs ? NodeDofs : fem::DofsFlag(fem::DofPosX | fem::DofPosY); }
  std::vector<r3::Vec> IsoparametricNodes() const override { std::vector<r3::Vec> n; for (auto &p : P) n.push_back({p[0], p[1], 0}); return n; }
  std::vector<double> Basis(r3::Vec v) const override { // quad8.go:51-63: the reference's expansion, term by term (same rounding)
    std::vector<double> N(8);
    const double x = v.X, y = v.Y;
    N[0] = 0.25 * (1 - x) * (1 - y) * (-1 - x - y);
    N[1] = 0.25 * (1 + x) * (1 - y) * (-1 + x - y);
    N[2] = 0.25 * (1 + x) * (1 + y) * (-1 + x + y);
    N[3] = 0.25 * (1 - x) * (1 + y) * (-1 - x + y);
    N[4] = 0.5 * (1 - x) * (1 + x) * (1 - y);
    N[5] = 0.5 * (1 + x) * (1 + y) * (1 - y);
    N[6] = 0.5 * (1 - x) * (1 + x) * (1 + y);
    N[7] = 0.5 * (1 - x) * (1 - y) * (1 + y);
    return N;
  }
  std::vector<double> BasisDiff(r3::Vec v) const override { // quad8.go:66-88: the reference's expansion, term by term (same rounding)
    std::vector<double> d(16);
    const double x = v.X, y = v.Y;
    d[0] = -(x/4-1./4)*(y-1) - ((y-1)*(x+y+1))/4;
    d[1] = ((y-1)*(y-x+1))/4 - (x/4+1./4)*(y-1);
    d[2] = (x/4+1./4)*(y+1) + ((y+1)*(x+y-1))/4;
    d[3] = (x/4-1./4)*(y+1) + ((y+1)*(x-y+1))/4;
    d[4] = ((x+1)*(y-1))/2 + (x/2-1./2)*(y-1);
    d[5] = -((y - 1) * (y + 1)) / 2;
    d[6] = -((x+1)*(y+1))/2 - (x/2-1./2)*(y+1);
    d[7] = ((y - 1) * (y + 1)) / 2;
    d[8] = -(x/4-1./4)*(y-1) - (x/4-1./4)*(x+y+1);
    d[9] = (x/4+1./4)*(y-x+1) + (x/4+1./4)*(y-1);
    d[10] = (x/4+1./4)*(y+1) + (x/4+1./4)*(x+y-1);
    d[11] = (x/4-1./4)*(x-y+1) - (x/4-1./4)*(y+1);
    d[12] = (x/2 - 1./2) * (x + 1);
    d[13] = -(x/2+1./2)*(y-1) - (x/2+1./2)*(y+1);
    d[14] = -(x/2 - 1./2) * (x + 1);
    d[15] = (x/2-1./2)*(y-1) + (x/2-1./2)*(y+1);
    return d;
  }
  int order() const { return QuadratureOrder <= 0 ? 3 : QuadratureOrder; } // quad8.go:98-104
  void Quadrature(std::vector<r3::Vec> &p, std::vector<double> &w) const override { if (!uniformGaussQuad2d(order(), order(), p, w)) { p.clear(); w.clear(); } }
  std::string String() const override { return "QUAD8(order=" + std::to_string(order()) + ")"; }
};

struct Triangle3 : IsoBase {
  int LenNodes() const override { return 3; }
  fem::DofsFlag Dofs() const override { return NodeDofs ? NodeDofs : fem::DofsFlag(fem::DofPosX | fem::DofPosY); }
  std::vector<r3::Vec> IsoparametricNodes() const override { return {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}}; }
  std::vector<double> Basis(r3::Vec v) const override { return {1 - v.X - v.Y, v.X, v.Y}; }
  std::vector<double> BasisDiff(r3::Vec) const override { return {-1, 1, 0, -1, 0, 1}; }
  int order() const { return QuadratureOrder <= 0 ? 1 : QuadratureOrder; } // triangles.go:70-76
  void Quadrature(std::vector<r3::Vec> &p, std::vector<double> &w) const override { if (!triangleQuads(order(), p, w)) { p.clear(); w.clear(); } }
  std::string String() const override { return "TRI6(order=" + std::to_string(order()) + ")"; } // sic, triangles.go:78
};

// Triangle6 is unusable in the reference (BasisDiff returns 8 values for 6 nodes, triangles.go:127-139 =>
// mat.NewDense panics at assembler_isoparametric.go:70).  The mirror keeps the type and the panic.
struct Triangle6 : IsoBase {
  int LenNodes() const override { return 6; }
  fem::DofsFlag Dofs() const override { return fem::DofsFlag(fem::DofPosX | fem::DofPosY); }
  std::vector<r3::Vec> IsoparametricNodes() const override { return {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {.5, .5, 0}, {0, .5, 0}, {.5, 0, 0}}; }
  std::vector<double> Basis(r3::Vec) const override { return std::vector<double>(6, 0.0); }
  std::vector<double> BasisDiff(r3::Vec) const override { return std::vector<double>(8, 0.0); } // length 8: d = 8/6 = 1
  int order() const { return QuadratureOrder <= 0 ? 2 : QuadratureOrder; }
  void Quadrature(std::vector<r3::Vec> &p, std::vector<double> &w) const override { if (!triangleQuads(order(), p, w)) { p.clear(); w.clear(); } }
  std::string String() const override { return "TRI6(order=" + std::to_string(order()) + ")"; }
};

} // namespace elements

namespace solids {

// solids.go:8-20 with the strain func replaced by a device B-kind tag
struct isoconstituter : fem::IsoConstituter {
  std::vector<double> C; int dim = 0; int bkind = -1; Error err;
  Error Constitutive(std::vector<double> &out, int &r, int &c) const override { out = C; r = c = dim; return err; }
  int BKind() const override { return bkind; }
};

// isotropic.go:13-102
struct Isotropic {
  double E = 0, Poisson = 0;
  double ShearModulus() const { return E / (2 + 2 * Poisson); }
  double lameParam2() const { return E * Poisson / ((1 + Poisson) * (1 - 2 * Poisson)); }
  Error Constitutive(std::vector<double> &C, int &r, int &c) const {
    const double G = ShearModulus(), l = lameParam2();
    r = c = 6;
    if (std::isnan(l) || std::isinf(l)) {
      char buf[96];
      std::snprintf(buf, sizeof buf, "matrix singular or near-singular with condition number %.4e", (1 + Poisson) * (1 - 2 * Poisson));
      C.clear();
      return Error::New(buf); // mat.Condition(...).Error()
    }
    C.assign(36, 0.0);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[i * 6 + j] = (i == j) ? l + 2 * G : l;
    C[21] = C[28] = C[35] = G;
    return Error::Nil();
  }
  isoconstituter Solid3D() const { isoconstituter k; int r, c; k.err = Constitutive(k.C, r, c); k.dim = 6; k.bkind = FEMB200_B_XYZ; return k; }
  isoconstituter PlaneStess() const { // sic, isotropic.go:55
    const double f = E / (1 - Poisson * Poisson);
    isoconstituter k; k.dim = 3; k.bkind = FEMB200_B_PLANE;
    k.C = {f, f * Poisson, 0, f * Poisson, f, 0, 0, 0, f * (1 - Poisson) / 2};
    return k;
  }
  isoconstituter PlaneStrain() const {
    const double f = E / (1 + Poisson) / (1 - 2 * Poisson), nuf = Poisson * f;
    isoconstituter k; k.dim = 3; k.bkind = FEMB200_B_PLANE;
    k.C = {f * (1 - Poisson), nuf, nuf, nuf, f * (1 - Poisson), nuf, nuf, nuf, f * (1 - 2 * Poisson)};
    return k;
  }
  isoconstituter Axisymmetric() const {
    const double nu = Poisson;
    double f = nu / (1 - nu), g = (1 - 2 * nu) / (2 * (1 - nu));
    const double factor = E * (1 - nu) / ((1 + nu) * (1 - 2 * nu));
    f *= factor; g *= factor;
    isoconstituter k; k.dim = 4; k.bkind = FEMB200_B_AXISYM;
    k.C = {factor, f, f, 0, f, factor, f, 0, f, f, factor, 0, 0, 0, 0, g};
    return k;
  }
};

// thermal.go:8-58
struct IsotropicConductivity {
  double K = 0;
  isoconstituter Solid3D() const { isoconstituter k; k.dim = 3; k.bkind = FEMB200_B_THERMAL; k.C = {K, 0, 0, 0, K, 0, 0, 0, K}; return k; }
  isoconstituter Plane() const { isoconstituter k; k.dim = 2; k.bkind = FEMB200_B_THERMAL; k.C = {K, 0, 0, K}; return k; }
  isoconstituter Axisymmetric() const { isoconstituter k; k.dim = 2; k.bkind = FEMB200_B_THERMAL_AXISYM; k.C = {K, 0, 0, K}; return k; }
};

} // namespace solids

namespace fem {

using GetElement = std::function<void(int64_t i, std::vector<int64_t> &elem, r3::Vec &xC, r3::Vec &yC)>;

// assembler.go:13-37 -- the assembler object; K lives on the GPU.
class GeneralAssembler {
public:
  // fem.NewGeneralAssembler (assembler.go:21-28).  `device` selects the B200.
  GeneralAssembler(const std::vector<r3::Vec> &nodes, DofsFlag modelDofs, int device = 0) : dofs_(modelDofs), n_nodes_((int64_t)nodes.size()) {
    int rc = femb200_ctx_create(device, &ctx_);
    if (rc) { init_err_ = "femb200: no usable sm_100 CUDA device (the GPU path has no CPU fallback)"; return; }
    static_assert(sizeof(r3::Vec) == 24, "r3.Vec must be 3 packed float64");
    rc = femb200_assembler_create(ctx_, nodes.empty() ? nullptr : &nodes[0].X, n_nodes_, Count(modelDofs) > 0 ? Count(modelDofs) : 1, 0, &ga_);
    if (rc) { init_err_ = std::string("femb200_assembler_create failed: ") + femb200_last_cuda_error(ctx_); return; }
    ksolid_.bind(ga_, n_nodes_ * Count(modelDofs));
  }
  ~GeneralAssembler() { femb200_assembler_destroy(ga_); femb200_ctx_destroy(ctx_); }
  GeneralAssembler(const GeneralAssembler &) = delete;
  GeneralAssembler &operator=(const GeneralAssembler &) = delete;

  lap::Sparse *Ksolid() { return &ksolid_; }                                  // assembler.go:31
  int64_t TotalDofs() const { return n_nodes_ * Count(dofs_); }               // assembler.go:34-37
  femb200_assembler *Device() { return ga_; }
  femb200_ctx *Context() { return ctx_; }

  // (*GeneralAssembler).AddIsoparametric (assembler_isoparametric.go:28-123)
  Error AddIsoparametric(const Isoparametric &elemT, const IsoConstituter &c, int64_t Nelem, const GetElement &getElement) {
    Prologue p;
    Error err = prologue(elemT, c, p);
    if (err) return err;
    std::vector<int64_t> conn;
    err = flatten(p, Nelem, getElement, conn);
    if (err) return err;
    int rc = femb200_add_isoparametric(ga_, &p.desc, conn.data(), 1, 0, Nelem);
    ksolid_.invalidate();
    return device_error(rc);
  }
  // additive fast path: connectivity already flat, no per-element callback
  Error AddIsoparametricFlat(const Isoparametric &elemT, const IsoConstituter &c, int64_t Nelem, const int64_t *conn) {
    Prologue p;
    Error err = prologue(elemT, c, p);
    if (err) return err;
    int rc = femb200_add_isoparametric(ga_, &p.desc, conn, 1, 0, Nelem);
    ksolid_.invalidate();
    return device_error(rc);
  }
  // (*GeneralAssembler).IsoparametricStrains (assembler_isoparametric.go:126-213)
  Error IsoparametricStrains(const std::vector<double> &displacements, const Isoparametric &elemT, const IsoConstituter &c, int64_t Nelem,
                             const GetElement &getElement, const std::function<void(int64_t, const double *, int)> &strainCallback) {
    if ((int64_t)displacements.size() != TotalDofs())
      return Error::New("displacements vector length " + std::to_string(displacements.size()) + " does not match total number of dofs " + std::to_string(TotalDofs()));
    Prologue p;
    Error err = prologue(elemT, c, p);
    if (err) return err;
    std::vector<int64_t> conn;
    err = flatten(p, Nelem, getElement, conn);
    if (err) return err;
    const int per = p.desc.nqp * p.desc.dimC;
    std::vector<double> strains((size_t)Nelem * per);
    int rc = femb200_isoparametric_strains(ga_, &p.desc, conn.data(), 1, 0, Nelem, displacements.data(), strains.data());
    err = device_error(rc);
    if (err) return err;
    for (int64_t e = 0; e < Nelem; ++e) strainCallback(e, strains.data() + e * per, per);
    return Error::Nil();
  }

  const std::string &InitError() const { return init_err_; }

private:
  struct Prologue {
    femb200_elem_desc desc{};
    std::vector<double> Npg, dNpg, wpg, C;
    std::vector<int32_t> dofmap;
  };
  // the once-per-call part of AddIsoparametric (assembler_isoparametric.go:29-71) + fem.go:120-139
  Error prologue(const Isoparametric &elemT, const IsoConstituter &c, Prologue &p) {
    if (!init_err_.empty()) return Error::New(init_err_);
    int rows = 0, cols = 0;
    Error err = c.Constitutive(p.C, rows, cols);                                   // :29
    if (err) return err;
    if (rows != cols) return Error::New("expected constitutive matrix to be square, got " + std::to_string(rows) + "x" + std::to_string(cols)); // :35
    const int nn = elemT.LenNodes();
    const std::vector<double> bd0 = elemT.BasisDiff(r3::Vec{});
    const int d = nn > 0 ? (int)bd0.size() / nn : 0;                               // :40
    const int ndn = Count(elemT.Dofs());                                           // :44
    std::vector<r3::Vec> upg;
    elemT.Quadrature(upg, p.wpg);                                                  // :56
    if (upg.empty() || upg.size() != p.wpg.size()) return Error::New("bad quadrature result from isoparametric element"); // :61
    if ((int)bd0.size() != d * nn)                                                 // mat.NewDense(d, nn, data) panics (:70)
      return Error::Panic("mat: bad matrix dimensions: BasisDiff returned " + std::to_string(bd0.size()) + " values for " + std::to_string(nn) + " nodes");
    for (const r3::Vec &pg : upg) {                                                // :66-71
      const std::vector<double> N = elemT.Basis(pg), dN = elemT.BasisDiff(pg);
      if ((int)N.size() != nn || (int)dN.size() != d * nn) return Error::Panic("mat: bad basis length from isoparametric element");
      p.Npg.insert(p.Npg.end(), N.begin(), N.end());
      p.dNpg.insert(p.dNpg.end(), dN.begin(), dN.end());
    }
    if (d < 1 || d > 3) return Error::New("dimsPerNode must be 1, 2 or 3");        // fem.go:123-125
    std::vector<int> dm;
    err = mapdofs(dofs_, elemT.Dofs(), dm);                                        // fem.go:127
    if (err) return err;
    if (Count(dofs_) != ndn) return Error::Panic("bad length");                    // fem.go:138 + assembler.go:148-150
    if (c.BKind() < 0) return Error::New("femb200: user-defined IsoConstituter cannot run on the GPU (no CPU fallback)");
    p.dofmap.assign(dm.begin(), dm.end());
    p.desc.d = d; p.desc.nn = nn; p.desc.ndn = ndn; p.desc.nqp = (int)upg.size(); p.desc.dimC = rows; p.desc.bkind = c.BKind();
    p.desc.model_dofs_per_node = Count(dofs_);
    p.desc.Npg = p.Npg.data(); p.desc.dNpg = p.dNpg.data(); p.desc.wpg = p.wpg.data(); p.desc.C = p.C.data(); p.desc.dofmap = p.dofmap.data();
    return Error::Nil();
  }
  // fem.go:140-144 + assembler_isoparametric.go:81-88: one callback per element, flattened for the device
  Error flatten(const Prologue &p, int64_t Nelem, const GetElement &getElement, std::vector<int64_t> &conn) {
    if (!getElement) return Error::Panic("nil argument to AddIsoparametric2");     // fem.go:121-122
    const int nn = p.desc.nn;
    conn.resize((size_t)Nelem * nn);
    std::vector<int64_t> elem;
    for (int64_t i = 0; i < Nelem; ++i) {
      r3::Vec x, y;
      elem.clear();
      getElement(i, elem, x, y);
      if ((int)elem.size() != nn)
        return Error::New("element #" + std::to_string(i) + " of " + std::to_string(elem.size()) + " nodes expected to be of " + std::to_string(nn) + " nodes"); // fem.go:143
      if (x != r3::Vec{} || y != r3::Vec{}) return Error::New("arbitrary constitutive orientation not implemented yet"); // :87
      for (int a = 0; a < nn; ++a) conn[(size_t)i * nn + a] = elem[a];
    }
    return Error::Nil();
  }
  // the reference's message texts (assembler_isoparametric.go:96,98,102,106)
  Error device_error(int rc) {
    if (rc == 0) return Error::Nil();
    int32_t kind = 0, qp = -1; int64_t e = -1;
    femb200_last_error(ga_, &kind, &e, &qp);
    const std::string es = std::to_string(e);
    switch (rc) {
    case FEMB200_ERR_NEG_DET: return Error::New("negative determinant of jacobian of element #" + es + ", Check node ordering");
    case FEMB200_ERR_ZERO_DET: return Error::New("zero determinant of jacobian of element #" + es + ", Check element shape for bad aspect ratio");
    case FEMB200_ERR_SOLVE: return Error::New("error calculating element #" + es + " form factor: matrix singular or near-singular");
    case FEMB200_ERR_NAN_SCALE: return Error::New("NaN scale value returned by SetStrainDisplacementMatrix at element #" + es + ", quad " + std::to_string(qp));
    case FEMB200_ERR_NODE_OOB: return Error::Panic("runtime error: index out of range (element #" + es + ")");
    case FEMB200_ERR_UNSUPPORTED: return Error::New("femb200: no compiled kernel for this element/constituter shape");
    case FEMB200_ERR_CUDA: return Error::New(std::string("femb200: CUDA error: ") + femb200_last_cuda_error(ctx_));
    default: return Error::New("femb200: status " + std::to_string(rc));
    }
  }

  femb200_ctx *ctx_ = nullptr;
  femb200_assembler *ga_ = nullptr;
  lap::Sparse ksolid_;
  DofsFlag dofs_;
  int64_t n_nodes_;
  std::string init_err_;
};

// fem.go:175-232 -- boundary-condition index helper (host side, used right after assembly)
class Fixity {
public:
  Fixity(DofsFlag modelDofs, int64_t numNodes) : nodes_((size_t)numNodes, 0), modelDofs_(modelDofs) {}
  void Fix(int64_t nodeIdx, DofsFlag fixed) { nodes_[(size_t)nodeIdx] |= (fixed & modelDofs_); }
  void Free(int64_t nodeIdx, DofsFlag freed) { nodes_[(size_t)nodeIdx] &= ~(freed & modelDofs_); }
  std::vector<int64_t> FreeDofs() const { return fixity(false); }
  std::vector<int64_t> FixedDofs() const { return fixity(true); }
private:
  std::vector<int64_t> fixity(bool getFixed) const {
    const int dofsPerNode = Count(modelDofs_);
    std::vector<int> dofIdx;
    for (int i = 0; i < 16; ++i) if (modelDofs_ & (1 << i)) dofIdx.push_back(i);
    std::vector<int64_t> out;
    for (size_t i = 0; i < nodes_.size(); ++i)
      for (size_t j = 0; j < dofIdx.size(); ++j) {
        const bool isFixed = (nodes_[i] & (1 << dofIdx[j])) != 0;
        if (getFixed == isFixed) out.push_back((int64_t)i * dofsPerNode + (int64_t)j);
      }
    return out;
  }
  std::vector<DofsFlag> nodes_;
  DofsFlag modelDofs_;
};

} // namespace fem
} // namespace gofem
