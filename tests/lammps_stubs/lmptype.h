// Minimal stand-ins for the LAMMPS (27May2021) declarations the shim classes use, so that
// `g++ -fsyntax-only` can check lammps-user-conp2_b200/shim/*.cpp against include/conp_b200.h on a
// machine without a LAMMPS tree.  Test infrastructure only: signatures, no behaviour.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#define FLERR __FILE__, __LINE__
#ifndef MAX
#define MIN(A, B) ((A) < (B) ? (A) : (B))
#define MAX(A, B) ((A) > (B) ? (A) : (B))
#endif
namespace fmt {  // LAMMPS bundles {fmt}; only the call shape is needed here
template <class... Args> std::string format(const char *, Args &&...) { return std::string(); }
}  // namespace fmt
namespace LAMMPS_NS {
typedef int tagint;
typedef int64_t bigint;
typedef int imageint;
}  // namespace LAMMPS_NS
