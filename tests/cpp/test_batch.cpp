// C++ host test of the C ABI through cpp/gemlab.hpp: the reference's doc-test of integ::mat_10_bdb
// (src/integ/mat_10_bdb.rs:60-119: Tet4, E = 480, nu = 1/3, 12x12 matrix printed with {:.0}) plus its error test
// (src/integ/mat_10_bdb.rs:401-409, "axisymmetric requires space_ndim = 2"), run on the GPU.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../cpp/gemlab.hpp"

using namespace gemlab;

int main() {
    Mesh mesh;
    mesh.ndim = 3;
    const double xx[4][3] = {{2, 3, 4}, {6, 3, 2}, {2, 5, 1}, {4, 3, 6}};
    for (size_t i = 0; i < 4; i++) mesh.points.push_back(Point{i, 0, {xx[i][0], xx[i][1], xx[i][2]}});
    // two cells so that the batch has more than one element; the second repeats the first
    mesh.cells.push_back(Cell{0, 1, GeoKind::Tet4, {0, 1, 2, 3}});
    mesh.cells.push_back(Cell{1, 1, GeoKind::Tet4, {0, 1, 2, 3}});
    static const double correct[12][12] = {
        {745, 540, 120, -5, 30, 60, -270, -240, 0, -470, -330, -180},
        {540, 1720, 270, -120, 520, 210, -120, -1080, -60, -300, -1160, -420},
        {120, 270, 565, 0, 150, 175, 0, -120, -270, -120, -300, -470},
        {-5, -120, 0, 145, -90, -60, -90, 120, 0, -50, 90, 60},
        {30, 520, 150, -90, 220, 90, 60, -360, -60, 0, -380, -180},
        {60, 210, 175, -60, 90, 145, 0, -120, -90, 0, -180, -230},
        {-270, -120, 0, -90, 60, 0, 180, 0, 0, 180, 60, 0},
        {-240, -1080, -120, 120, -360, -120, 0, 720, 0, 120, 720, 240},
        {0, -60, -270, 0, -60, -90, 0, 0, 180, 0, 120, 180},
        {-470, -300, -120, -50, 0, 0, 180, 120, 0, 340, 180, 120},
        {-330, -1160, -300, 90, -380, -180, 60, 720, 120, 180, 820, 360},
        {-180, -420, -470, 60, -180, -230, 0, 240, 180, 120, 360, 520}};
    try {
        Context ctx(0);
        DeviceMesh dmesh = ctx.upload(mesh);
        integ::CommonArgs args;
        auto dd = integ::Coef::uniform(integ::lin_elasticity(480.0, 1.0 / 3.0, false));
        auto kk = integ::batch::mat_10_bdb(ctx, dmesh, {0, 2}, args, dd);
        if (kk.size() != 2 * 144) {
            std::printf("FAIL: size %zu\n", kk.size());
            return 1;
        }
        double worst = 0;
        for (int e = 0; e < 2; e++)
            for (int i = 0; i < 12; i++)
                for (int j = 0; j < 12; j++) worst = std::fmax(worst, std::fabs(kk[e * 144 + i + 12 * j] - correct[i][j]));
        if (worst > 1e-10) {
            std::printf("FAIL: max abs difference %.3e\n", worst);
            return 1;
        }
        args.axisymmetric = true;
        bool threw = false;
        try {
            integ::batch::mat_10_bdb(ctx, dmesh, {0, 2}, args, dd);
        } catch (const StrError &e) {
            threw = std::strcmp(e.what(), "axisymmetric requires space_ndim = 2") == 0;
        }
        if (!threw) {
            std::printf("FAIL: axisymmetric error string\n");
            return 1;
        }
        std::printf("OK max abs difference %.3e\n", worst);
    } catch (const StrError &e) {
        std::printf("FAIL: %s\n", e.what());
        return 1;
    }
    return 0;
}
