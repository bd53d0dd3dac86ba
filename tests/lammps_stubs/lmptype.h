// Minimal single-rank stand-in for the LAMMPS (27May2021) classes the shim uses: enough behaviour for
// tests/lammps_stubs/mock_lammps_main.cpp to drive lammps-user-conp2_b200/shim/*.cpp through the hook
// order of a real run (no LAMMPS tree exists on this machine), and for `g++ -fsyntax-only` to check the
// shim against include/conp_b200.h.  Test infrastructure only.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#define FLERR __FILE__, __LINE__
#ifndef MAX
#define MIN(A, B) ((A) < (B) ? (A) : (B))
#define MAX(A, B) ((A) > (B) ? (A) : (B))
#endif
typedef double FFT_SCALAR;
namespace fmt {  // LAMMPS bundles {fmt}; the shim only formats one double with {:.8g}
inline std::string format(const char *f, double v) {
  std::string s(f);
  const size_t at = s.find("{:.8g}");
  if (at == std::string::npos) return s;
  char buf[64];
  snprintf(buf, sizeof(buf), "%.8g", v);
  return s.substr(0, at) + buf + s.substr(at + 6);
}
}  // namespace fmt
namespace LAMMPS_NS {
typedef int tagint;
typedef int64_t bigint;
typedef int imageint;
}  // namespace LAMMPS_NS
