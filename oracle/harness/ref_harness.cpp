// ref_harness -- TEST INFRASTRUCTURE ONLY. Drives the UNMODIFIED reference (niclasberg/biblbm) library,
// compiled from /root/reference by oracle/build_ref.sh, through exactly the per-step call sequence its
// drivers use (progs/rbc_in_shear/rbc_in_shear.cpp:154-186, progs/rbcs_in_shear/rbcs_in_shear.cpp:165-193):
//     computeVelocity -> setExternalVector -> applyProcessingFunctional(wrap_ibm_dynamics3D) -> collideAndStream
// and dumps raw state so that the C restatement (oracle/biblbm_oracle.c) and the CUDA path can be compared
// with it. This file is our own harness code: it contains no reference source, only calls into its API.
//
// Sub-commands (all file formats are raw little-endian, documented at the readers/writers in tests/kit.py and tests/golden/make_golden.py):
//   ref_harness mesh2bin <in.msh> <out.mesh>        gmsh -> {i64 nv, i64 nt, f64 xyz[nv*3], i64 tri[nt*3]}
//   ref_harness shape    <in.mesh> <out_prefix>     connectivity as built by ParticleShape::compute_connectivity
//   ref_harness deflate  <in.mesh> <out.f64> G K_area K_bend vol_frac N_it     shrink_rbc_volume (setup-time host loop)
//   ref_harness run      <config.txt>               time loop with dumps / timers
#include "palabos3D.h"
#include "palabos3D.hh"
#include "fsi/headers.h"
#include "fsi/headers.hh"
#include "fsi/GuoBoundaryCondition.h"
#include "fsi/GuoBoundaryCondition.hh"
#include "common/calibrate_rbc_shape.h"
#include "common/shear_flow_setup.h"
#include "common/particle_position_regulator.h"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

using namespace plb;
using namespace fsi;

#define DESCRIPTOR descriptors::ForcedPhaseD3Q19Descriptor
typedef long long i64;

// ------------------------------------------------------------------------------------------------ io helpers
static void write_f64(const std::string& fname, const std::vector<double>& v) {
    FILE* f = std::fopen(fname.c_str(), "wb");
    if (!f) { std::fprintf(stderr, "cannot write %s\n", fname.c_str()); std::exit(2); }
    if (!v.empty()) std::fwrite(&v[0], sizeof(double), v.size(), f);
    std::fclose(f);
}
static void write_i64(const std::string& fname, const std::vector<i64>& v) {
    FILE* f = std::fopen(fname.c_str(), "wb");
    if (!f) { std::fprintf(stderr, "cannot write %s\n", fname.c_str()); std::exit(2); }
    if (!v.empty()) std::fwrite(&v[0], sizeof(i64), v.size(), f);
    std::fclose(f);
}
static std::vector<double> read_f64(const std::string& fname) {
    FILE* f = std::fopen(fname.c_str(), "rb");
    if (!f) { std::fprintf(stderr, "cannot read %s\n", fname.c_str()); std::exit(2); }
    std::fseek(f, 0, SEEK_END); long n = std::ftell(f); std::fseek(f, 0, SEEK_SET);
    std::vector<double> v(n / sizeof(double));
    if (!v.empty() && std::fread(&v[0], sizeof(double), v.size(), f) != v.size()) std::exit(2);
    std::fclose(f);
    return v;
}
struct MeshBin {
    i64 nv, nt;
    std::vector<Array<double, 3> > verts;
    std::vector<unsigned int> tris;
};
static MeshBin read_mesh_bin(const std::string& fname) {
    MeshBin m;
    FILE* f = std::fopen(fname.c_str(), "rb");
    if (!f) { std::fprintf(stderr, "cannot read %s\n", fname.c_str()); std::exit(2); }
    if (std::fread(&m.nv, 8, 1, f) != 1 || std::fread(&m.nt, 8, 1, f) != 1) std::exit(2);
    m.verts.resize(m.nv);
    for (i64 i = 0; i < m.nv; ++i) {
        double p[3];
        if (std::fread(p, 8, 3, f) != 3) std::exit(2);
        m.verts[i] = Array<double, 3>(p[0], p[1], p[2]);
    }
    m.tris.resize(3 * m.nt);
    for (i64 i = 0; i < 3 * m.nt; ++i) {
        i64 t;
        if (std::fread(&t, 8, 1, f) != 1) std::exit(2);
        m.tris[i] = (unsigned int)t;
    }
    std::fclose(f);
    return m;
}
static void store_mesh(ParticleShapeLibrary<double>& lib, MeshBin& m, const char* tag) {
    lib.store_mesh(&m.verts[0], &m.tris[0], (unsigned int)m.nv, (unsigned int)m.nt, tag);
}

// ------------------------------------------------------------------------------------------------ mesh2bin
static int cmd_mesh2bin(int argc, char** argv) {
    if (argc < 4) return 1;
    ParticleShapeLibrary<double> lib;
    lib.read_and_store_mesh(argv[2], "S");
    const ParticleShape<double>* s = lib.get_by_tag("S");
    FILE* f = std::fopen(argv[3], "wb");
    i64 nv = s->count_vertices(), nt = s->count_triangles();
    std::fwrite(&nv, 8, 1, f); std::fwrite(&nt, 8, 1, f);
    for (i64 i = 0; i < nv; ++i) { double p[3] = { s->get_vertex(i)[0], s->get_vertex(i)[1], s->get_vertex(i)[2] }; std::fwrite(p, 8, 3, f); }
    for (i64 i = 0; i < nt; ++i) { i64 t[3] = { s->get_triangle(i).i0, s->get_triangle(i).i1, s->get_triangle(i).i2 }; std::fwrite(t, 8, 3, f); }
    std::fclose(f);
    std::printf("nv=%lld nt=%lld\n", nv, nt);
    return 0;
}

// ------------------------------------------------------------------------------------------------ shape
static int cmd_shape(int argc, char** argv) {
    if (argc < 4) return 1;
    MeshBin m = read_mesh_bin(argv[2]);
    ParticleShapeLibrary<double> lib;
    store_mesh(lib, m, "S");
    const ParticleShape<double>* s = lib.get_by_tag("S");
    std::string p(argv[3]);
    std::vector<i64> li; std::vector<double> lf;
    for (ParticleShape<double>::link_const_iterator it = s->links_begin(); it != s->links_end(); ++it) {
        li.push_back(it->i0); li.push_back(it->i1); li.push_back(it->t1); li.push_back(it->t2); lf.push_back(it->length);
    }
    write_i64(p + ".links.i64", li); write_f64(p + ".links.f64", lf);
    std::vector<i64> ti; std::vector<double> tf;
    for (ParticleShape<double>::triangle_const_iterator it = s->triangles_begin(); it != s->triangles_end(); ++it) {
        ti.push_back(it->i0); ti.push_back(it->i1); ti.push_back(it->i2); tf.push_back(it->area); tf.push_back(it->l_mean);
    }
    write_i64(p + ".tris.i64", ti); write_f64(p + ".tris.f64", tf);
    std::vector<i64> pi; std::vector<double> pf;
    for (ParticleShape<double>::triangle_pair_const_iterator it = s->triangle_pairs_begin(); it != s->triangle_pairs_end(); ++it) {
        pi.push_back(it->i0); pi.push_back(it->i1); pi.push_back(it->i2); pi.push_back(it->i3);
        pf.push_back(it->cos_theta0); pf.push_back(it->sin_theta0);
    }
    write_i64(p + ".pairs.i64", pi); write_f64(p + ".pairs.f64", pf);
    std::vector<double> sc;
    sc.push_back(s->get_area()); sc.push_back(s->get_volume()); sc.push_back(s->get_radius());
    sc.push_back(s->get_center()[0]); sc.push_back(s->get_center()[1]); sc.push_back(s->get_center()[2]);
    write_f64(p + ".scalars.f64", sc);
    std::printf("links=%ld tris=%ld pairs=%ld area=%.17g volume=%.17g radius=%.17g\n",
                (long)lf.size(), (long)(ti.size() / 3), (long)(pi.size() / 4), sc[0], sc[1], sc[2]);
    return 0;
}

static void pack_params(const RBCParameters<double>& p, std::vector<double>& out) {
    out.push_back(p.shear_modulus); out.push_back(p.L0); out.push_back(p.k_area_global); out.push_back(p.k_area_local);
    out.push_back(p.k_volume); out.push_back(p.vol_desired); out.push_back(p.k_bend); out.push_back(p.theta0);
}
static RBCParameters<double> unpack_params(const double* v) {
    RBCParameters<double> p;
    p.shear_modulus = v[0]; p.L0 = v[1]; p.k_area_global = v[2]; p.k_area_local = v[3];
    p.k_volume = v[4]; p.vol_desired = v[5]; p.k_bend = v[6]; p.theta0 = v[7];
    return p;
}

// ------------------------------------------------------------------------------------------------ deflate
// out.f64 = [8 params][nv*3 vertex positions][area, volume]
static int cmd_deflate(int argc, char** argv) {
    if (argc < 9) return 1;
    MeshBin m = read_mesh_bin(argv[2]);
    ParticleShapeLibrary<double> lib;
    store_mesh(lib, m, "S");
    const ParticleShape<double>* s = lib.get_by_tag("S");
    double G = std::atof(argv[4]), K_area = std::atof(argv[5]), K_bend = std::atof(argv[6]);
    double vol_frac = std::atof(argv[7]);
    int n_it = std::atoi(argv[8]);
    RBCParameters<double> params = create_rbc_params(s, G, K_area, K_bend);
    RBCParticle<double> rbc(s, params);
    if (n_it > 0) shrink_rbc_volume(rbc, vol_frac * s->get_volume(), n_it);
    std::vector<double> out;
    pack_params(rbc.params(), out);
    for (i64 i = 0; i < rbc.count_nodes(); ++i) for (int d = 0; d < 3; ++d) out.push_back(rbc.get_node(i).pos[d]);
    out.push_back(rbc.area()); out.push_back(rbc.volume());
    write_f64(argv[3], out);
    std::printf("deflated: area=%.17g volume=%.17g vol_desired=%.17g\n", rbc.area(), rbc.volume(), rbc.params().vol_desired);
    return 0;
}

// ------------------------------------------------------------------------------------------------ run
struct Config {
    std::map<std::string, std::string> kv;
    bool has(const std::string& k) const { return kv.find(k) != kv.end(); }
    std::string str(const std::string& k, const std::string& d = "") const { return has(k) ? kv.find(k)->second : d; }
    double num(const std::string& k, double d) const { return has(k) ? std::atof(kv.find(k)->second.c_str()) : d; }
    i64 integer(const std::string& k, i64 d) const { return has(k) ? std::atoll(kv.find(k)->second.c_str()) : d; }
};
static Config read_config(const char* fname) {
    Config c; std::ifstream in(fname); std::string line;
    if (!in) { std::fprintf(stderr, "cannot read config %s\n", fname); std::exit(2); }
    while (std::getline(in, line)) {
        size_t eq = line.find('=');
        if (eq == std::string::npos || line[0] == '#') continue;
        c.kv[line.substr(0, eq)] = line.substr(eq + 1);
    }
    return c;
}

template <class T>
struct SineInit {
    SineInit(i64 nx_, i64 ny_, i64 nz_, T u0_) : nx(nx_), ny(ny_), nz(nz_), u0(u0_) {}
    void operator()(plint iX, plint iY, plint iZ, T& rho, Array<T, 3>& u) const {
        rho = (T)1 + 0.01 * std::cos(2. * M_PI * (T)iX / (T)nx) * std::cos(2. * M_PI * (T)iZ / (T)nz);
        u[0] = u0 * std::sin(2. * M_PI * (T)iY / (T)ny);
        u[1] = u0 * std::sin(2. * M_PI * (T)iZ / (T)nz);
        u[2] = u0 * std::sin(2. * M_PI * (T)iX / (T)nx);
    }
    i64 nx, ny, nz; T u0;
};

// In the MPI build MultiBlockLattice3D::get() hands out a proxy cell (parallelism/parallelMultiBlockLattice3D.hh:69-87),
// so raw cell data is reached through the single atomic block instead: local index = global index - getLocation().
static plb::Cell<double, DESCRIPTOR>& raw_cell(MultiBlockLattice3D<double, DESCRIPTOR>& lattice, plint x, plint y, plint z) {
    const plint id = lattice.getLocalInfo().getBlocks()[0];
    BlockLattice3D<double, DESCRIPTOR>& b = lattice.getComponent(id);
    const Dot3D loc = b.getLocation();
    return b.get(x - loc.x, y - loc.y, z - loc.z);
}

static void dump_lattice(MultiBlockLattice3D<double, DESCRIPTOR>& lattice, const std::string& fname) {
    // Palabos serialised-cell layout: 19 fBar populations then the 4 external scalars (src/core/cell.hh:103-121),
    // cells in [x][y][z] order, z fastest (src/atomicBlock/blockLattice3D.hh:671-689).
    const plint nx = lattice.getNx(), ny = lattice.getNy(), nz = lattice.getNz();
    std::vector<double> v((size_t)nx * ny * nz * 23);
    size_t o = 0;
    for (plint x = 0; x < nx; ++x) for (plint y = 0; y < ny; ++y) for (plint z = 0; z < nz; ++z) {
        plb::Cell<double, DESCRIPTOR>& c = raw_cell(lattice, x, y, z);
        for (int i = 0; i < 19; ++i) v[o++] = c[i];
        for (int i = 0; i < 4; ++i) v[o++] = *c.getExternal(i);
    }
    write_f64(fname, v);
}
// debugging aid: the whole atomic block including its envelope, local index order
static void dump_block(MultiBlockLattice3D<double, DESCRIPTOR>& lattice, const std::string& fname) {
    const plint id = lattice.getLocalInfo().getBlocks()[0];
    BlockLattice3D<double, DESCRIPTOR>& b = lattice.getComponent(id);
    std::vector<double> v((size_t)b.getNx() * b.getNy() * b.getNz() * 23);
    size_t o = 0;
    for (plint x = 0; x < b.getNx(); ++x) for (plint y = 0; y < b.getNy(); ++y) for (plint z = 0; z < b.getNz(); ++z) {
        plb::Cell<double, DESCRIPTOR>& c = b.get(x, y, z);
        for (int i = 0; i < 19; ++i) v[o++] = c[i];
        for (int i = 0; i < 4; ++i) v[o++] = *c.getExternal(i);
    }
    write_f64(fname, v);
}
static void dump_velocity(MultiTensorField3D<double, 3>& vel, const std::string& fname) {
    const plint nx = vel.getNx(), ny = vel.getNy(), nz = vel.getNz();
    std::vector<double> v((size_t)nx * ny * nz * 3);
    size_t o = 0;
    for (plint x = 0; x < nx; ++x) for (plint y = 0; y < ny; ++y) for (plint z = 0; z < nz; ++z) {
        const plint id = vel.getLocalInfo().getBlocks()[0];
        TensorField3D<double, 3>& b = vel.getComponent(id);
        const Dot3D loc = b.getLocation();
        Array<double, 3> const& u = b.get(x - loc.x, y - loc.y, z - loc.z);
        v[o++] = u[0]; v[o++] = u[1]; v[o++] = u[2];
    }
    write_f64(fname, v);
}
template <class FSI>
static void dump_particles(FSI& fsi, i64 np, const std::string& vert_name, const std::string& part_name) {
    std::vector<double> vv, pp;
    for (i64 i = 0; i < np; ++i) {
        const ParticleBase3D<double>* p;
        if (!fsi.get_particle(i, p)) { std::fprintf(stderr, "particle %lld missing\n", i); std::exit(3); }
        for (plint n = 0; n < p->count_nodes(); ++n) {
            const Vertex<double>& v = p->get_node(n);
            for (int d = 0; d < 3; ++d) vv.push_back(v.pos[d]);
            for (int d = 0; d < 3; ++d) vv.push_back(v.vel[d]);
            for (int d = 0; d < 3; ++d) vv.push_back(v.force[d]);
        }
        pp.push_back(p->area()); pp.push_back(p->volume());
        for (int d = 0; d < 3; ++d) pp.push_back(p->center_of_mass()[d]);
        geo::Rect<double> bb = p->bounding_box();
        pp.push_back(bb.x0); pp.push_back(bb.x1); pp.push_back(bb.y0); pp.push_back(bb.y1); pp.push_back(bb.z0); pp.push_back(bb.z1);
    }
    write_f64(vert_name, vv); write_f64(part_name, pp);
    // rigid particles: com, velocity, angular velocity, orientation, force, torque (19 doubles each)
    std::vector<double> rr;
    for (i64 i = 0; i < np; ++i) {
        const ParticleBase3D<double>* p;
        fsi.get_particle(i, p);
        const RigidParticle3D<double>* r = dynamic_cast<const RigidParticle3D<double>*>(p);
        if (!r) continue;
        for (int d = 0; d < 3; ++d) rr.push_back(r->center_of_mass()[d]);
        for (int d = 0; d < 3; ++d) rr.push_back(r->velocity()[d]);
        for (int d = 0; d < 3; ++d) rr.push_back(r->angular_velocity()[d]);
        for (int d = 0; d < 4; ++d) rr.push_back(r->orientation()[d]);
        for (int d = 0; d < 3; ++d) rr.push_back(r->force()[d]);
        for (int d = 0; d < 3; ++d) rr.push_back(r->torque()[d]);
    }
    if (!rr.empty()) { std::string n = vert_name; n.replace(n.find("vert_"), 5, "rigid_"); write_f64(n, rr); }
}

template <class Periodicity>
static int run(const Config& cfg) {
    const i64 nx = cfg.integer("nx", 16), ny = cfg.integer("ny", 16), nz = cfg.integer("nz", 16);
    const i64 steps = cfg.integer("steps", 1), dump_every = cfg.integer("dump_every", 0);
    const std::string out = cfg.str("out", ".");
    const double omega = cfg.num("omega", 1.0), omega_inner = cfg.num("omega_inner", 1.0 / 3.0);
    const Array<double, 3> body(cfg.num("body_x", 0), cfg.num("body_y", 0), cfg.num("body_z", 0));
    const bool shear = cfg.str("walls", "none") == "shear";
    const bool lbm_only = cfg.integer("lbm_only", 0) != 0;   // only collideAndStream each step (externals frozen)
    const bool timing = cfg.integer("timing", 0) != 0;
    const double regulator_p = cfg.num("regulator_p", 0.0);

    MultiBlockLattice3D<double, DESCRIPTOR> lattice(nx, ny, nz, new GuoExternalForceVOFDynamics<double, DESCRIPTOR>(omega, omega_inner));
    Box3D domain = lattice.getBoundingBox();
    lattice.toggleInternalStatistics(false);

    OnLatticeBoundaryCondition3D<double, DESCRIPTOR>* bc = createInterpBoundaryCondition3D<double, DESCRIPTOR>();
    if (shear) {
        shearFlowSetup(lattice, cfg.num("u_wall", 0.01), *bc);   // x,z periodic, velocity planes at y=0 and y=ny-1
    } else {
        lattice.periodicity().toggle(0, Periodicity::get_x());
        lattice.periodicity().toggle(1, Periodicity::get_y());
        lattice.periodicity().toggle(2, Periodicity::get_z());
        if (cfg.str("init", "sine") == "uniform")
            initializeAtEquilibrium(lattice, domain, cfg.num("rho", 1.0), Array<double, 3>(cfg.num("ux", 0), cfg.num("uy", 0), cfg.num("uz", 0)));
        else
            initializeAtEquilibrium(lattice, domain, SineInit<double>(nx, ny, nz, cfg.num("u0", 0.05)));
        lattice.initialize();
    }

    // rbc_in_pipe.cpp:285-343 (setup_flow): Guo off-lattice wall condition on the pipe surface + NoDynamics outside of it
    PipeBoundary<double>* guo_pipe = 0;
    GuoRigidWallBoundary<double, DESCRIPTOR>* guo = 0;
    if (cfg.integer("guo_pipe", 0) != 0 && cfg.has("pipe_radius")) {
        guo_pipe = new PipeBoundary<double>(cfg.num("pipe_y0", 0.5 * (ny - 1)), cfg.num("pipe_z0", 0.5 * (nz - 1)), cfg.num("pipe_radius", 0), cfg.num("pipe_length", nx - 1));
        guo = new GuoRigidWallBoundary<double, DESCRIPTOR>(*guo_pipe, lattice);
        guo->insert();
        defineDynamics<double, DESCRIPTOR>(lattice, lattice.getBoundingBox(), new BoundaryOutside<double>(*guo_pipe, std::sqrt(2)),
                                           new plb::NoDynamics<double, DESCRIPTOR>());
        if (cfg.has("guo_wall_ux")) guo->set_wall_velocity(Array<double, 3>(cfg.num("guo_wall_ux", 0), cfg.num("guo_wall_uy", 0), cfg.num("guo_wall_uz", 0)));
        lattice.initialize();
        if (cfg.has("guo_dump")) {
            std::vector<double> nd;
            for (GuoRigidWallBoundary<double, DESCRIPTOR>::iterator it = guo->begin(); it != guo->end(); ++it) {
                nd.push_back(it->location.x); nd.push_back(it->location.y); nd.push_back(it->location.z);
                nd.push_back(it->direction.x); nd.push_back(it->direction.y); nd.push_back(it->direction.z);
                nd.push_back(it->delta);
            }
            write_f64(out + "/guo_nodes.f64", nd);
        }
    }

    // Optional frozen synthetic external field (force + phase), read from a file in serialised-cell order [x][y][z][4]
    if (cfg.has("externals")) {
        std::vector<double> ext = read_f64(cfg.str("externals"));
        size_t o = 0;
        for (plint x = 0; x < nx; ++x) for (plint y = 0; y < ny; ++y) for (plint z = 0; z < nz; ++z) {
            plb::Cell<double, DESCRIPTOR>& c = raw_cell(lattice, x, y, z);
            for (int i = 0; i < 4; ++i) *c.getExternal(i) = ext[o++];
        }
        lattice.duplicateOverlaps(modif::staticVariables);
    }
    // Optional initial populations override (same layout as the dumps)
    if (cfg.has("populations")) {
        std::vector<double> pop = read_f64(cfg.str("populations"));
        size_t o = 0;
        for (plint x = 0; x < nx; ++x) for (plint y = 0; y < ny; ++y) for (plint z = 0; z < nz; ++z) {
            plb::Cell<double, DESCRIPTOR>& c = raw_cell(lattice, x, y, z);
            for (int i = 0; i < 19; ++i) c[i] = pop[o++];
            for (int i = 0; i < 4; ++i) *c.getExternal(i) = pop[o++];
        }
        lattice.duplicateOverlaps(modif::staticVariables);
    }

    // Particles: one shape (mesh file), per-particle vertex positions, one RBC parameter set
    ParticleShapeLibrary<double> shape_library;
    MeshBin mesh; i64 np = 0;
    ImmersedBoundaryDynamics3D<double, DESCRIPTOR, Periodicity>* fsi = 0;
    ParticlePositionRegulator<double>* regulator = 0;
    if (cfg.has("mesh")) {
        mesh = read_mesh_bin(cfg.str("mesh"));
        store_mesh(shape_library, mesh, "RBC");
    }
    fsi = new ImmersedBoundaryDynamics3D<double, DESCRIPTOR, Periodicity>(lattice, shape_library);
    if (cfg.has("particles")) {
        // particles file: [8 RBC params][np * nv * 3 positions]
        std::vector<double> pf = read_f64(cfg.str("particles"));
        RBCParameters<double> params = unpack_params(&pf[0]);
        const i64 nv = mesh.nv;
        np = (i64)(pf.size() - 8) / (nv * 3);
        const ParticleShape<double>* shape = shape_library.get_by_tag("RBC");
        const bool capsule = cfg.str("particle_type", "rbc") == "capsule";     // params = {G, C, nu_s, kb, 0, 0, 0, 0}
        const bool rigid = cfg.str("particle_type", "rbc") == "rigid";  // params = {density, damping, relaxation_factor, scale}; rigid_state file: [np][13] com, vel, angvel, quat
        if (rigid) {
            std::vector<double> rsf = read_f64(cfg.str("rigid_state"));
            np = (i64)rsf.size() / 13;
            for (i64 p = 0; p < np; ++p) {
                const double* r = &rsf[p * 13];
                RigidParticle3D<double> rp(shape);
                rp.density() = pf[0]; rp.damping() = pf[1]; rp.relaxation_factor() = pf[2]; rp.scale() = pf[3];
                rp.orientation() = Quaternion<double>(r[9], r[10], r[11], r[12]);
                rp.update();
                rp.set_center_of_mass(Array<double, 3>(r[0], r[1], r[2]));
                rp.velocity() = Array<double, 3>(r[3], r[4], r[5]);
                rp.angular_velocity() = Array<double, 3>(r[6], r[7], r[8]);
                rp.update();
                fsi->add_particle(&rp);
            }
            std::vector<double> sh;
            Matrix<double, 3> I = shape->get_moment_of_inertia();
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) sh.push_back(I(a, b));
            sh.push_back(shape->get_volume());
            write_f64(out + "/rigid_shape.f64", sh);
        }
        const bool semirigid = cfg.str("particle_type", "rbc") == "semirigid";  // params = {l0, k_in_plane, k_out_of_plane, k_bend, 0, 0, 0, 0}
        for (i64 p = 0; p < (rigid ? 0 : np); ++p) {
            if (semirigid) {
                SemiRigidParticleParams<double> sp; sp.l0 = pf[0]; sp.k_in_plane = pf[1]; sp.k_out_of_plane = pf[2]; sp.k_bend = pf[3];
                SemiRigidParticle3D<double> plt(shape, sp);
                for (i64 n = 0; n < nv; ++n) for (int d = 0; d < 3; ++d) plt.get_node(n).pos[d] = pf[8 + (p * nv + n) * 3 + d];
                plt.update();
                fsi->add_particle(&plt);
            } else if (capsule) {
                CapsuleParameters<double> cp; cp.G = pf[0]; cp.C = pf[1]; cp.nu_s = pf[2]; cp.kb = pf[3];
                DeformableCapsuleParticle<double> cap(shape, cp);
                for (i64 n = 0; n < nv; ++n) for (int d = 0; d < 3; ++d) cap.get_node(n).pos[d] = pf[8 + (p * nv + n) * 3 + d];
                cap.update();
                fsi->add_particle(&cap);
            } else {
                RBCParticle<double> rbc(shape, params);
                for (i64 n = 0; n < nv; ++n) for (int d = 0; d < 3; ++d) rbc.get_node(n).pos[d] = pf[8 + (p * nv + n) * 3 + d];
                rbc.update();
                fsi->add_particle(&rbc);
            }
        }
        if (regulator_p != 0.0) {
            regulator = new ParticlePositionRegulator<double>(
                Array<double, 3>(cfg.num("regulator_x", 0.5 * (nx - 1)), cfg.num("regulator_y", 0.5 * (ny - 1)), cfg.num("regulator_z", 0.5 * (nz - 1))), regulator_p);
            fsi->add_force_decorator(regulator);
        }
    }
    Boundary<double>* boundary = 0;
    if (cfg.has("plate_y0")) {                                   // rbcs_in_shear.cpp:122-123
        boundary = new ParallelPlatesBoundary<double>(cfg.num("plate_y0", 0), cfg.num("plate_y1", ny - 1));
        fsi->set_boundary(boundary);
    } else if (cfg.has("pipe_radius")) {                         // rbc_in_pipe.cpp:42,304,733 (PipeBoundary(y0, z0, radius, length))
        boundary = new PipeBoundary<double>(cfg.num("pipe_y0", 0.5 * (ny - 1)), cfg.num("pipe_z0", 0.5 * (nz - 1)), cfg.num("pipe_radius", 0),
                                            cfg.num("pipe_length", nx - 1));
        fsi->set_boundary(boundary);
    }
    // rbc_in_pipe.cpp:150-178,736-741: wall interaction (ForceDecorator) and particle-particle interaction (PPInteractionForce)
    if (cfg.has("wall_morse") && boundary) {
        std::istringstream is(cfg.str("wall_morse")); double r_min, beta, strength; is >> r_min >> beta >> strength;
        fsi->add_force_decorator(new WallInteraction<double, MorsePotential<double> >(*boundary, MorsePotential<double>(r_min, beta, strength)));
    }
    if (cfg.has("wall_spring") && boundary) {
        std::istringstream is(cfg.str("wall_spring")); double r_min, strength; is >> r_min >> strength;
        fsi->add_force_decorator(new WallInteraction<double, SpringPotential<double> >(*boundary, SpringPotential<double>(r_min, strength)));
    }
    if (cfg.has("pp_morse")) {
        std::istringstream is(cfg.str("pp_morse")); double r_min, beta, strength; is >> r_min >> beta >> strength;
        fsi->add_pp_force(new PPPotentialForce<double, MorsePotential<double> >(MorsePotential<double>(r_min, beta, strength)));
    }
    if (cfg.has("pp_spring")) {
        std::istringstream is(cfg.str("pp_spring")); double r_min, strength; is >> r_min >> strength;
        fsi->add_pp_force(new PPPotentialForce<double, SpringPotential<double> >(SpringPotential<double>(r_min, strength)));
    }
    fsi->init();

    std::auto_ptr<MultiTensorField3D<double, 3> > velocity = generateMultiTensorField<double, 3>(domain);

    dump_lattice(lattice, out + "/pop_init.f64");
    if (cfg.integer("dump_block", 0)) dump_block(lattice, out + "/blk_init.f64");
    if (np > 0) dump_particles(*fsi, np, out + "/vert_init.f64", out + "/part_init.f64");

    double t_vel = 0, t_ext = 0, t_ib = 0, t_lbm = 0;
    const i64 warm = cfg.integer("warmup", 0);
    for (i64 it = 0; it < steps + warm; ++it) {
        const bool timed = it >= warm;
        double t0 = MPI_Wtime();
        if (!lbm_only) computeVelocity(lattice, *velocity, domain);
        double t1 = MPI_Wtime();
        if (!lbm_only) setExternalVector(lattice, domain, DESCRIPTOR<double>::ExternalField::forceBeginsAt, body);
        double t2 = MPI_Wtime();
        if (!lbm_only) applyProcessingFunctional(wrap_ibm_dynamics3D(*fsi), domain, lattice, *velocity);
        double t3 = MPI_Wtime();
        if (cfg.integer("debug_manual_cs", 0)) {
            // debugging aid (fully periodic only): run the atomic-block step by hand and dump the block before the envelope refresh
            const plint id = lattice.getLocalInfo().getBlocks()[0];
            BlockLattice3D<double, DESCRIPTOR>& b = lattice.getComponent(id);
            b.collideAndStream(b.getBoundingBox());
            std::ostringstream s; s << (it + 1);
            dump_block(lattice, out + "/pre_" + s.str() + ".f64");
            lattice.duplicateOverlaps(modif::staticVariables);
        } else
        lattice.collideAndStream();
        double t4 = MPI_Wtime();
        if (timed) { t_vel += t1 - t0; t_ext += t2 - t1; t_ib += t3 - t2; t_lbm += t4 - t3; }
        const i64 done = it + 1;
        if (dump_every > 0 && (done % dump_every == 0 || done == steps)) {
            std::ostringstream s; s << done;
            dump_lattice(lattice, out + "/pop_" + s.str() + ".f64");
            if (cfg.integer("dump_block", 0)) dump_block(lattice, out + "/blk_" + s.str() + ".f64");
            if (!lbm_only) dump_velocity(*velocity, out + "/vel_" + s.str() + ".f64");
            if (np > 0) dump_particles(*fsi, np, out + "/vert_" + s.str() + ".f64", out + "/part_" + s.str() + ".f64");
        }
    }
    if (timing) {
        std::printf("{\"steps\": %lld, \"cells\": %lld, \"vertices\": %lld, \"t_compute_velocity\": %.6f, \"t_set_external\": %.6f, \"t_ib\": %.6f, \"t_collide_stream\": %.6f}\n",
                    steps, nx * ny * nz, np * (i64)mesh.nv, t_vel, t_ext, t_ib, t_lbm);
        Profile::write_report((out + "/profile").c_str());
    }
    delete fsi;
    return 0;
}

static int cmd_run(int argc, char** argv) {
    if (argc < 3) return 1;
    Config cfg = read_config(argv[2]);
    std::string per = cfg.str("periodicity", "ppp");
    if (cfg.str("walls", "none") == "shear") per = "pfp";
    if (per == "ppp") return run<Periodicity3D<double, true, true, true> >(cfg);
    if (per == "pfp") return run<Periodicity3D<double, true, false, true> >(cfg);
    if (per == "pff") return run<Periodicity3D<double, true, false, false> >(cfg);
    std::fprintf(stderr, "unsupported periodicity %s\n", per.c_str());
    return 1;
}

int main(int argc, char** argv) {
    plbInit(&argc, &argv);
    if (argc < 2) { std::fprintf(stderr, "usage: ref_harness mesh2bin|shape|deflate|run ...\n"); return 1; }
    std::string cmd(argv[1]);
    int rc = 1;
    if (cmd == "mesh2bin") rc = cmd_mesh2bin(argc, argv);
    else if (cmd == "shape") rc = cmd_shape(argc, argv);
    else if (cmd == "deflate") rc = cmd_deflate(argc, argv);
    else if (cmd == "run") rc = cmd_run(argc, argv);
    if (rc == 1) std::fprintf(stderr, "bad arguments for %s\n", cmd.c_str());
    return rc;
}
