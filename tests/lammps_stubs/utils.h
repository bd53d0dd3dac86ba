#pragma once
#include "pointers.h"
#include <cstdlib>
#include <regex>
#include <string>
namespace LAMMPS_NS {
namespace utils {
inline double numeric(const char *, int, const char *s, bool, LAMMPS *) { return strtod(s, nullptr); }
inline int inumeric(const char *, int, const char *s, bool, LAMMPS *) { return (int)strtol(s, nullptr, 10); }
inline bigint bnumeric(const char *, int, const char *s, bool, LAMMPS *) { return strtoll(s, nullptr, 10); }
inline void bounds(const char *, int, const std::string &str, bigint nmin, bigint nmax, int &nlo, int &nhi, class Error *) {
  const size_t star = str.find('*');
  if (star == std::string::npos) { nlo = nhi = atoi(str.c_str()); return; }
  nlo = star == 0 ? (int)nmin : atoi(str.substr(0, star).c_str());
  nhi = star + 1 == str.size() ? (int)nmax : atoi(str.substr(star + 1).c_str());
}
inline void logmesg(LAMMPS *lmp, const std::string &m) { lmp->log += m; fputs(m.c_str(), stdout); }
inline bool strmatch(const std::string &text, const std::string &pattern) { return std::regex_search(text, std::regex(pattern)); }
inline char *strdup(const std::string &s) { char *r = new char[s.size() + 1]; strcpy(r, s.c_str()); return r; }
}  // namespace utils
}  // namespace LAMMPS_NS
