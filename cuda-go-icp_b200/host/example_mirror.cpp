// Compile check + usage example of the C++ mirror (host/goicp_b200.hpp): the reference's
// main.cpp:47-59 protocol, verbatim member names.
#include <cstdio>
#include <vector>
#include "goicp_b200.hpp"

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: %s <model.txt|ply> <data.txt|ply> [mse]\n", argv[0]); return 2; }
    float *m = nullptr, *d = nullptr; int nm = 0, nd = 0;
    if (goicp_load_cloud(argv[1], 0.1f, 1.0f, 1234, &m, &nm) || goicp_load_cloud(argv[2], 0.1f, 1.0f, 1235, &d, &nd)) { std::fprintf(stderr, "%s\n", goicp_io_last_error()); return 1; }
    try {
        goicp_b200::GoICP goicp(argc > 3 ? (float)std::atof(argv[3]) : 1e-3f);
        goicp.pModel = reinterpret_cast<goicp_b200::POINT3D*>(m); goicp.Nm = nm;
        goicp.pData = reinterpret_cast<goicp_b200::POINT3D*>(d); goicp.Nd = nd;
        goicp.BuildDT();
        goicp.Register();
        std::printf("optError %g  t = (%g, %g, %g)\n", goicp.optError, goicp.optT.val[0][0], goicp.optT.val[1][0], goicp.optT.val[2][0]);
    } catch (const std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
    goicp_free_cloud(m); goicp_free_cloud(d);
    return 0;
}
