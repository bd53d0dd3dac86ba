#pragma once
#include "lmptype.h"
#include <string>
namespace LAMMPS_NS {
class LAMMPS;
namespace utils {
double numeric(const char *, int, const char *, bool, LAMMPS *);
int inumeric(const char *, int, const char *, bool, LAMMPS *);
bigint bnumeric(const char *, int, const char *, bool, LAMMPS *);
void bounds(const char *, int, const std::string &, bigint, bigint, int &, int &, class Error *);
void logmesg(LAMMPS *, const std::string &);
bool strmatch(const std::string &, const std::string &);
char *strdup(const std::string &);
}  // namespace utils
}  // namespace LAMMPS_NS
