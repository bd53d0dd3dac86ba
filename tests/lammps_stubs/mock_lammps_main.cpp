// Mock LAMMPS host (single rank): drives the shim classes FixConpB200 / FixConqB200 / FixCondB200 and
// PPPMCONPB200 (lammps-user-conp2_b200/shim) through the hook order of Verlet::setup and Verlet::run
//     init -> setup_post_neighbor -> [kspace->setup() inside setup_pre_force] -> setup_pre_force
//     -> post_force -> [PPPM::compute: make_rho]  -> (move atoms) post_neighbor -> pre_force -> post_force
// against libconp_b200.so, on a system the Python tests wrote to a file, and writes what a LAMMPS user
// would see (atom->q of the electrode atoms, the fix scalar, forces, kspace energy, the density brick the
// force PPPM would transform, the log file) to another file.  Test infrastructure: no LAMMPS tree exists here.
//
//   mock_lammps <system.bin> <out.bin> fix-args...      (fix-args as in the deck: ID group style Nevery ...)
#include "atom.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "group.h"
#include "input.h"
#include "kspace.h"
#include "memory.h"
#include "pair.h"
#include "update.h"
#include "variable.h"

#include "fix_conp.h"
#include "pppm_conp.h"

#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

using namespace LAMMPS_NS;

namespace {
struct Blob {
  std::map<std::string, std::vector<double>> f;
  std::map<std::string, std::vector<int>> i;
};
// records: int32 name_len | name | int32 kind (0 int32, 1 float64) | int64 count | data
Blob read_blob(const char *path) {
  Blob b;
  FILE *fp = fopen(path, "rb");
  if (!fp) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  int nl;
  while (fread(&nl, 4, 1, fp) == 1) {
    std::string name(nl, ' ');
    int kind;
    long long cnt;
    if (fread(&name[0], 1, nl, fp) != (size_t)nl || fread(&kind, 4, 1, fp) != 1 || fread(&cnt, 8, 1, fp) != 1) exit(2);
    if (kind == 0) { b.i[name].resize(cnt); if (fread(b.i[name].data(), 4, cnt, fp) != (size_t)cnt) exit(2); }
    else { b.f[name].resize(cnt); if (fread(b.f[name].data(), 8, cnt, fp) != (size_t)cnt) exit(2); }
  }
  fclose(fp);
  return b;
}
void put(FILE *fp, const std::string &name, const std::vector<double> &v) {
  const int nl = (int)name.size(), kind = 1;
  const long long cnt = (long long)v.size();
  fwrite(&nl, 4, 1, fp); fwrite(name.data(), 1, nl, fp); fwrite(&kind, 4, 1, fp); fwrite(&cnt, 8, 1, fp);
  fwrite(v.data(), 8, v.size(), fp);
}
}  // namespace

int main(int argc, char **argv) {
  if (argc < 11) { fprintf(stderr, "usage: mock_lammps system.bin out.bin ID group style Nevery group2 eta value logfile [keywords]\n"); return 2; }
  Blob in = read_blob(argv[1]);
  LAMMPS lmp;
  lmp.memory = new Memory(&lmp);
  lmp.error = new Error(&lmp);
  lmp.comm = new Comm(&lmp);
  lmp.domain = new Domain(&lmp);
  lmp.group = new Group(&lmp);
  lmp.atom = new Atom(&lmp);
  lmp.update = new Update(&lmp);
  lmp.force = new Force(&lmp);
  lmp.input = new Input(&lmp);
  lmp.input->variable = new Variable(&lmp);
  Atom *atom = lmp.atom;
  Domain *domain = lmp.domain;
  Force *force = lmp.force;

  // ---- the system ---------------------------------------------------------------------------------
  const int n = in.i["natoms"][0], ntypes = in.i["ntypes"][0];
  atom->natoms = n; atom->nlocal = n; atom->nghost = 0; atom->nmax = n; atom->ntypes = ntypes;
  atom->tag = in.i["tag"].data(); atom->type = in.i["type"].data(); atom->mask = in.i["mask"].data();
  atom->q = in.f["q"].data();
  std::vector<double> &xs = in.f["x"], fs(3 * (size_t)n, 0.0);
  std::vector<double *> xrows(n), frows(n);
  for (int i = 0; i < n; ++i) { xrows[i] = xs.data() + 3 * (size_t)i; frows[i] = fs.data() + 3 * (size_t)i; }
  atom->x = xrows.data(); atom->f = frows.data();
  for (int a = 0; a < 3; ++a) {
    domain->boxlo[a] = in.f["boxlo"][a]; domain->prd[a] = in.f["prd"][a]; domain->boxhi[a] = domain->boxlo[a] + domain->prd[a];
    domain->periodicity[a] = in.i["periodic"][a];
  }
  domain->xprd = domain->prd[0]; domain->yprd = domain->prd[1]; domain->zprd = domain->prd[2];
  domain->xperiodic = domain->periodicity[0]; domain->yperiodic = domain->periodicity[1]; domain->zperiodic = domain->periodicity[2];
  lmp.group->add("all"); lmp.group->add("eleleft"); lmp.group->add("eleright");   // mask bits 1, 2, 4
  // pair style with a coulomb cut-off
  Pair *pair = new Pair(&lmp);
  force->pair = pair; force->pair_style = "lj/cut/coul/long";
  pair->cut_coul = in.f["cut_coul"][0];
  std::vector<double> &cs = in.f["cutsq"];
  std::vector<double *> csrows(ntypes + 1);
  for (int t = 0; t <= ntypes; ++t) csrows[t] = cs.data() + (size_t)t * (ntypes + 1);
  pair->cutsq = csrows.data();
  // kspace style: pppm/conp when the deck asks for it, else a plain PPPM (only g_ewald / accuracy are read)
  const bool conp_kspace = in.i["kspace_is_conp"][0] != 0;
  PPPM *pppm = conp_kspace ? new PPPMCONPB200(&lmp) : new PPPM(&lmp);
  force->kspace = pppm;
  pppm->g_ewald = in.f["g_ewald"][0]; pppm->accuracy = in.f["accuracy"][0];
  pppm->slabflag = in.i["slabflag"][0]; pppm->slab_volfactor = in.f["slab_volfactor"][0];
  if (in.i.count("mesh"))
    pppm->mock_tables(in.i["mesh"].data(), in.i["order"][0], in.f["rho_coeff"].data(), in.f["greensfn"].data(),
                      in.f["shift"][0], in.f["shiftone"][0]);
  if (in.f.count("variable_value")) { lmp.input->variable->name = "dv"; lmp.input->variable->value = in.f["variable_value"][0]; }

  FILE *out = fopen(argv[2], "wb");
  int rc = 0;
  try {
    // ---- fix ID group style ... ------------------------------------------------------------------
    const std::string style = argv[5];
    FixConpB200 *fix = style == "conq" ? new FixConqB200(&lmp, argc - 3, argv + 3)
                     : style == "cond" ? new FixCondB200(&lmp, argc - 3, argv + 3) : new FixConpB200(&lmp, argc - 3, argv + 3);
    if (in.i.count("fix_modify")) {   // fix_modify ID ehgo coeff <type> eta auto  (tests/il_onelayer/input:104-106)
      char a0[] = "ehgo", a1[] = "coeff", a4[] = "auto";
      std::string ty = std::to_string(in.i["fix_modify"][0]), eta = std::to_string(in.f["fix_modify_eta"][0]);
      char *margs[5] = {a0, a1, &ty[0], &eta[0], a4};
      fix->modify_param(5, margs);
    }
    const int nsteps = in.f.count("x2") ? 1 : 0;
    lmp.update->ntimestep = 0; lmp.update->laststep = nsteps;
    auto dump = [&](const std::string &tag) {
      std::vector<double> q(atom->q, atom->q + n), f(fs), sc(1, fix->compute_scalar()), en(1, pppm->energy), ec(1, pair->eng_coul);
      put(out, "q" + tag, q); put(out, "f" + tag, f); put(out, "scalar" + tag, sc); put(out, "kspace_energy" + tag, en);
      put(out, "eng_coul" + tag, ec);
      if (conp_kspace) {   // what the force PPPM would transform: PPPM::compute starts with particle_map + make_rho
        pppm->mock_make_rho();
        std::vector<double> rho((size_t)pppm->nx_pppm * pppm->ny_pppm * pppm->nz_pppm);
        size_t k = 0;
        for (int iz = 0; iz < pppm->nz_pppm; ++iz)
          for (int iy = 0; iy < pppm->ny_pppm; ++iy)
            for (int ix = 0; ix < pppm->nx_pppm; ++ix) rho[k++] = pppm->mock_density(ix, iy, iz);
        put(out, "density" + tag, rho);
        // ghost layers must be zero: LAMMPS' reverse_comm adds them onto the owners
        std::vector<double> ghost(1, pppm->mock_density(-1, 0, 0) + pppm->mock_density(0, -1, 0) + pppm->mock_density(0, 0, -1));
        put(out, "ghost" + tag, ghost);
      }
    };
    // ---- Verlet::setup ---------------------------------------------------------------------------
    fix->init();
    fix->setup_post_neighbor();
    fix->setup_pre_force(0);            // calls force->kspace->setup() itself (reference fix_conp.cpp:387-391)
    pppm->energy = 1.0; pair->eng_coul = 0.0;   // kspace energy must be non-zero for the self term to be added (reference fix_conp.cpp:1166)
    fix->post_force(0);
    dump("0");
    // ---- one Verlet::run step with moved atoms ----------------------------------------------------
    if (nsteps) {
      xs = in.f["x2"];
      for (int i = 0; i < n; ++i) xrows[i] = xs.data() + 3 * (size_t)i;
      std::fill(fs.begin(), fs.end(), 0.0);
      lmp.update->ntimestep = 1;
      fix->post_neighbor();
      fix->pre_force(0);
      pppm->energy = 1.0; pair->eng_coul = 0.0;   // kspace energy must be non-zero for the self term to be added (reference fix_conp.cpp:1166)
      fix->post_force(0);
      fix->end_of_step();
      dump("1");
    }
    put(out, "kspace_setups", std::vector<double>(1, (double)pppm->setups));
    delete fix;   // closes the log file
  } catch (const std::exception &e) {
    fprintf(stderr, "ERROR: %s\n", e.what());
    const std::string m = e.what();
    put(out, "error", std::vector<double>(m.begin(), m.end()));
    rc = 1;
  }
  fclose(out);
  return rc;
}
