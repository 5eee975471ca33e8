// C++ host test of the device-resident and multi-device paths of cpp/gemlab.hpp: DeviceSlab (gl_integ_slab), MultiContext
// (gl_multi_*), the device lattice generator (gl_mesh_block), the MSH text reader (gl_mesh_from_text) and gl_integ_fused.
// Everything is compared with the plain host-output call on the same cells: the results must be bit-identical.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../cpp/gemlab.hpp"

using namespace gemlab;

static const char *TWO_CUBES =
    "3 12 2 0 0\n"
    "0 0 0.0 0.0 0.0\n1 0 1.0 0.0 0.0\n2 0 1.0 1.0 0.0\n3 0 0.0 1.0 0.0\n4 0 0.0 0.0 1.0\n5 0 1.0 0.0 1.0\n"
    "6 0 1.0 1.0 1.0\n7 0 0.0 1.0 1.0\n8 0 0.0 0.0 2.0\n9 0 1.0 0.0 2.0\n10 0 1.0 1.0 2.0\n11 0 0.0 1.0 2.0\n"
    "0 1 hex8 0 1 2 3 4 5 6 7\n1 2 hex8 4 5 6 7 8 9 10 11\n";

static Mesh two_cubes() {
    Mesh m;
    m.ndim = 3;
    const double x[12][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}, {0, 0, 2}, {1, 0, 2}, {1, 1, 2}, {0, 1, 2}};
    for (size_t i = 0; i < 12; i++) m.points.push_back(Point{i, 0, {x[i][0], x[i][1], x[i][2]}});
    m.cells.push_back(Cell{0, 1, GeoKind::Hex8, {0, 1, 2, 3, 4, 5, 6, 7}});
    m.cells.push_back(Cell{1, 2, GeoKind::Hex8, {4, 5, 6, 7, 8, 9, 10, 11}});
    return m;
}

#define REQUIRE(cond, what)                  \
    if (!(cond)) {                           \
        std::printf("FAIL: %s\n", what);     \
        return 1;                            \
    }

int main() {
    try {
        Context ctx(0);
        Mesh mesh = two_cubes();
        DeviceMesh dmesh = ctx.upload(mesh);
        integ::CommonArgs args;
        auto dd = integ::Coef::uniform(integ::lin_elasticity(480.0, 1.0 / 3.0, false));
        const auto ref = integ::batch::mat_10_bdb(ctx, dmesh, {0, 2}, args, dd);

        // 1. device-resident output
        DeviceSlab slab = integ::batch::mat_10_bdb_device(ctx, dmesh, {0, 2}, args, dd);
        ctx.sync();
        REQUIRE(slab.size() == ref.size() && slab.cells().second == 2, "slab extent");
        REQUIRE(slab.download(ctx.handle()) == ref, "slab values");

        // 2. the MSH text of the same mesh through the library's parser
        DeviceMesh tmesh = ctx.mesh_from_text(TWO_CUBES);
        REQUIRE(integ::batch::mat_10_bdb(ctx, tmesh, {0, 2}, args, dd) == ref, "mesh_from_text");
        bool threw = false;
        try {
            ctx.mesh_from_text("3 1 1 0 0\n0 0 0.0 0.0\n");
        } catch (const StrError &e) {
            threw = std::strcmp(e.what(), "cannot read point z coordinate") == 0;
        }
        REQUIRE(threw, "reader error string");

        // 3. Block::new_cube(2).set_ndiv([1,1,2]) on the device == the two cubes scaled; its first cell is the unit cube scaled by 2
        DeviceMesh bmesh = ctx.block(GeoKind::Hex8, {1, 1, 2}, {0, 0, 0}, {1, 1, 2});
        REQUIRE(integ::batch::mat_10_bdb(ctx, bmesh, {0, 2}, args, dd) == ref, "device block");

        // 4. several contexts on device 0 behind the multi-device driver: concatenation is bit-identical
        MultiContext multi({0, 0});
        multi.upload(mesh);
        REQUIRE(multi.integ(GL_OP_MAT_10_BDB, {0, 2}, args, dd) == ref, "multi host output");
        auto slabs = multi.integ_slabs(GL_OP_MAT_10_BDB, {0, 2}, args, dd);
        REQUIRE(slabs.size() == 2 && slabs[0].cells().second == 1 && slabs[1].cells().first == 1, "multi pieces");
        auto a = slabs[0].download(multi.ctx(0)), b = slabs[1].download(multi.ctx(1));
        a.insert(a.end(), b.begin(), b.end());
        REQUIRE(a == ref, "multi slabs");

        // 5. mass matrix + source vector from one geometry pass
        auto rho = integ::Coef::uniform({2.7}), src = integ::Coef::uniform({9.81});
        auto both = integ::batch::fused(ctx, dmesh, {0, 2}, args, {{GL_OP_MAT_01_NSN, &rho}, {GL_OP_VEC_01_NS, &src}});
        REQUIRE(both[0] == integ::batch::mat_01_nsn(ctx, dmesh, {0, 2}, args, rho), "fused matrix");
        REQUIRE(both[1] == integ::batch::vec_01_ns(ctx, dmesh, {0, 2}, args, src), "fused vector");
        std::printf("OK\n");
    } catch (const StrError &e) {
        std::printf("FAIL: %s\n", e.what());
        return 1;
    }
    return 0;
}
