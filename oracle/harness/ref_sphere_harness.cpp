// ref_sphere_harness -- TEST INFRASTRUCTURE ONLY. Miniature of progs/sphere_in_ext_flow on the UNMODIFIED reference.
//
// The driver's source file is included IN PLACE from /root/reference (its main() renamed by the preprocessor, nothing is copied
// into this repository): that brings in the driver-local MultiDirectForcingFunctional3D (sphere_in_ext_flow.cpp:14-120), its
// DESCRIPTOR (ForcedD3Q19Descriptor), its Periodicity (false, false, true) and progs/common/extensional_flow_setup.h. This
// harness then repeats the driver's set-up (sphere_in_ext_flow.cpp:135-215) at a configurable size and runs its time loop
// (sphere_in_ext_flow.cpp:285-292):
//     applyProcessingFunctional(new MultiDirectForcingFunctional3D(fsi, relaxIterations), domain, lattice, velocity);
//     lattice.collideAndStream();
// dumping raw state after every step.
//
//   ref_sphere_harness <config.txt>
// config keys: geo_nx, geo_ny, geo_nz, tau, gamma, relaxation_factor, relax_iterations, density, damping, scale, pos_x, pos_y, pos_z,
//              steps, mesh (binary mesh, see ref_harness.cpp), out
#define main sphere_in_ext_flow_driver_main
#include "sphere_in_ext_flow/sphere_in_ext_flow.cpp"
#undef main

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

typedef long long i64;

static void write_f64(const std::string& fname, const std::vector<double>& v) {
    FILE* f = std::fopen(fname.c_str(), "wb");
    if (!f) { std::fprintf(stderr, "cannot write %s\n", fname.c_str()); std::exit(2); }
    if (!v.empty()) std::fwrite(&v[0], sizeof(double), v.size(), f);
    std::fclose(f);
}
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
struct MeshBin { i64 nv, nt; std::vector<Array<double, 3> > verts; std::vector<unsigned int> tris; };
static MeshBin read_mesh_bin(const std::string& fname) {
    MeshBin m;
    FILE* f = std::fopen(fname.c_str(), "rb");
    if (!f) { std::fprintf(stderr, "cannot read %s\n", fname.c_str()); std::exit(2); }
    if (std::fread(&m.nv, 8, 1, f) != 1 || std::fread(&m.nt, 8, 1, f) != 1) std::exit(2);
    m.verts.resize(m.nv);
    for (i64 i = 0; i < m.nv; ++i) { double p[3]; if (std::fread(p, 8, 3, f) != 3) std::exit(2); m.verts[i] = Array<double, 3>(p[0], p[1], p[2]); }
    m.tris.resize(3 * m.nt);
    for (i64 i = 0; i < 3 * m.nt; ++i) { i64 t; if (std::fread(&t, 8, 1, f) != 1) std::exit(2); m.tris[i] = (unsigned int)t; }
    std::fclose(f);
    return m;
}

// serialised-cell layout of ForcedD3Q19Descriptor: 19 fBar populations + 3 external scalars (the force)
static void dump_lattice(MultiBlockLattice3D<double, DESCRIPTOR>& lattice, const std::string& fname) {
    const plint nx = lattice.getNx(), ny = lattice.getNy(), nz = lattice.getNz();
    const plint id = lattice.getLocalInfo().getBlocks()[0];
    BlockLattice3D<double, DESCRIPTOR>& b = lattice.getComponent(id);
    const Dot3D loc = b.getLocation();
    std::vector<double> v((size_t)nx * ny * nz * 22);
    size_t o = 0;
    for (plint x = 0; x < nx; ++x) for (plint y = 0; y < ny; ++y) for (plint z = 0; z < nz; ++z) {
        plb::Cell<double, DESCRIPTOR>& c = b.get(x - loc.x, y - loc.y, z - loc.z);
        for (int i = 0; i < 19; ++i) v[o++] = c[i];
        for (int i = 0; i < 3; ++i) v[o++] = *c.getExternal(i);
    }
    write_f64(fname, v);
}
// stored boundary velocities as computeVelocity reports them on boundary cells, the whole lattice ([x][y][z][3]; bulk cells report
// their fluid velocity, so only the boundary cells of this dump are used)
static void dump_velocity(MultiTensorField3D<double, 3>& vel, const std::string& fname) {
    const plint nx = vel.getNx(), ny = vel.getNy(), nz = vel.getNz();
    const plint id = vel.getLocalInfo().getBlocks()[0];
    TensorField3D<double, 3>& b = vel.getComponent(id);
    const Dot3D loc = b.getLocation();
    std::vector<double> v((size_t)nx * ny * nz * 3);
    size_t o = 0;
    for (plint x = 0; x < nx; ++x) for (plint y = 0; y < ny; ++y) for (plint z = 0; z < nz; ++z) {
        Array<double, 3> const& u = b.get(x - loc.x, y - loc.y, z - loc.z);
        v[o++] = u[0]; v[o++] = u[1]; v[o++] = u[2];
    }
    write_f64(fname, v);
}
template <class FSI>
static void dump_particle(FSI& fsi, const std::string& out, const std::string& tag) {
    ParticleBase3D<double>* p;
    if (!fsi.get_particle(0, p)) { std::fprintf(stderr, "particle missing\n"); std::exit(3); }
    std::vector<double> vv, rr, pp;
    for (plint n = 0; n < p->count_nodes(); ++n) {
        const Vertex<double>& v = p->get_node(n);
        for (int d = 0; d < 3; ++d) vv.push_back(v.pos[d]);
        for (int d = 0; d < 3; ++d) vv.push_back(v.vel[d]);
        for (int d = 0; d < 3; ++d) vv.push_back(v.force[d]);
    }
    RigidParticle3D<double>* r = dynamic_cast<RigidParticle3D<double>*>(p);
    for (int d = 0; d < 3; ++d) rr.push_back(r->center_of_mass()[d]);
    for (int d = 0; d < 3; ++d) rr.push_back(r->velocity()[d]);
    for (int d = 0; d < 3; ++d) rr.push_back(r->angular_velocity()[d]);
    for (int d = 0; d < 4; ++d) rr.push_back(r->orientation()[d]);
    for (int d = 0; d < 3; ++d) rr.push_back(r->force()[d]);
    for (int d = 0; d < 3; ++d) rr.push_back(r->torque()[d]);
    pp.push_back(p->area()); pp.push_back(p->volume());
    for (int d = 0; d < 3; ++d) pp.push_back(p->center_of_mass()[d]);
    geo::Rect<double> bb = p->bounding_box();
    pp.push_back(bb.x0); pp.push_back(bb.x1); pp.push_back(bb.y0); pp.push_back(bb.y1); pp.push_back(bb.z0); pp.push_back(bb.z1);
    write_f64(out + "/vert_" + tag + ".f64", vv);
    write_f64(out + "/rigid_" + tag + ".f64", rr);
    write_f64(out + "/part_" + tag + ".f64", pp);
}

int main(int argc, char** argv) {
    plbInit(&argc, &argv);
    if (argc < 2) { std::fprintf(stderr, "usage: ref_sphere_harness <config.txt>\n"); return 1; }
    Config cfg = read_config(argv[1]);
    const plint geoNx = cfg.integer("geo_nx", 24), geoNy = cfg.integer("geo_ny", 24), geoNz = cfg.integer("geo_nz", 15);
    const i64 steps = cfg.integer("steps", 1);
    const std::string out = cfg.str("out", ".");
    const double tau = cfg.num("tau", 0.55), omega = 1. / tau;
    const double gamma_lb = cfg.num("gamma", 1e-3);
    const plint relaxIterations = cfg.integer("relax_iterations", 4);

    ParticleShapeLibrary<double> shape_library;
    MeshBin mesh = read_mesh_bin(cfg.str("mesh"));
    shape_library.store_mesh(&mesh.verts[0], &mesh.tris[0], (unsigned int)mesh.nv, (unsigned int)mesh.nt, "sphere");
    ParticleShape<double>* shape = shape_library.get_by_tag("sphere");

    // sphere_in_ext_flow.cpp:172-179 (+ scale and damping, which the driver leaves at their defaults)
    Array<double, 3> particlePosition(cfg.num("pos_x", 0.5 * geoNx), cfg.num("pos_y", 0.8 * geoNy), cfg.num("pos_z", 0.5 * geoNz));
    RigidParticle3D<double> particle(shape);
    particle.relaxation_factor() = cfg.num("relaxation_factor", 0.3);
    particle.density() = cfg.num("density", 1.0);
    particle.damping() = cfg.num("damping", 0.0);
    particle.scale() = cfg.num("scale", 1.0);
    particle.update();
    particle.set_center_of_mass(particlePosition);
    particle.velocity() = extensionalFlowVelocity(particlePosition, geoNx, geoNy, gamma_lb);
    particle.update();

    // sphere_in_ext_flow.cpp:195-215
    MultiBlockLattice3D<double, DESCRIPTOR> lattice(geoNx + 1, geoNy + 1, geoNz + 1, new GuoExternalForceBGKdynamics<double, DESCRIPTOR>(omega));
    Box3D domain = lattice.getBoundingBox();
    lattice.toggleInternalStatistics(false);
    OnLatticeBoundaryCondition3D<double, DESCRIPTOR>* boundaryCondition = createInterpBoundaryCondition3D<double, DESCRIPTOR>();
    extensionalFlowSetup(lattice, gamma_lb, *boundaryCondition);
    ImmersedBoundaryDynamics3D<double, DESCRIPTOR, Periodicity> fsi(lattice, shape_library);
    std::auto_ptr<MultiTensorField3D<double, 3> > velocity = generateMultiTensorField<double, 3>(domain);
    fsi.add_particle(&particle);
    fsi.init();

    std::vector<double> sh;
    Matrix<double, 3> I = shape->get_moment_of_inertia();
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) sh.push_back(I(a, b));
    sh.push_back(shape->get_volume());
    write_f64(out + "/rigid_shape.f64", sh);

    dump_lattice(lattice, out + "/pop_init.f64");
    computeVelocity(lattice, *velocity, domain);
    dump_velocity(*velocity, out + "/bvel.f64");
    dump_particle(fsi, out, "init");
    for (i64 it = 0; it < steps; ++it) {
        applyProcessingFunctional(new MultiDirectForcingFunctional3D<double, DESCRIPTOR, Periodicity>(fsi, relaxIterations), domain, lattice, *velocity);
        std::ostringstream s; s << (it + 1);
        if (cfg.integer("dump_mid", 0)) dump_lattice(lattice, out + "/mid_" + s.str() + ".f64");
        lattice.collideAndStream();
        dump_lattice(lattice, out + "/pop_" + s.str() + ".f64");
        dump_particle(fsi, out, s.str());
    }
    return 0;
}
