// placeholder CLI (filled in with the TOML driver)
#include "goicp_b200.h"
#include <cstdio>
int main(int argc, char** argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: %s <config.toml>\n", argv[0]); return 2; }
    goicp_result r;
    int rc = goicp_run_toml(argv[1], 1234, 1235, &r);
    if (rc) { std::fprintf(stderr, "goicp_run_toml failed: %d\n", rc); return 1; }
    return 0;
}
