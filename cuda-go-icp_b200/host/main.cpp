// goicp_b200_cli -- headless replacement of the reference's `cis5650_fgo_icp <config.toml>`
// (src/main.cpp:14-28): same TOML file, no window; prints the result as one JSON line and writes
// [io].output / [io].visualization if they are set.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "goicp_b200.h"

int main(int argc, char** argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: %s <config.toml> [seed_model=1234] [seed_data=1235]\n", argv[0]); return 2; }
    const unsigned sm = argc > 2 ? (unsigned)std::atoi(argv[2]) : 1234u, sd = argc > 3 ? (unsigned)std::atoi(argv[3]) : 1235u;
    goicp_result r;
    const int rc = goicp_run_toml(argv[1], sm, sd, &r);
    if (rc) { std::fprintf(stderr, "error %d: %s\n", rc, goicp_io_last_error()); return 1; }
    static const char* paths[] = {"none", "certified", "early_sse_below_thresh", "queue_empty", "cancelled"};
    std::printf("{\"R\": [%.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g], \"t\": [%.9g, %.9g, %.9g], \"sse\": %.9g, \"sse_thresh\": %.9g, "
                "\"lower_bound\": %.9g, \"exit_path\": \"%s\", \"rot_pops\": %lld, \"trans_pops\": %lld, \"bound_evals\": %lld, \"seconds\": %.6f}\n",
                r.R[0], r.R[1], r.R[2], r.R[3], r.R[4], r.R[5], r.R[6], r.R[7], r.R[8], r.t[0], r.t[1], r.t[2], r.sse, r.sse_thresh, r.best_lb,
                paths[r.exit_path], (long long)r.rot_pops, (long long)r.trans_pops, (long long)r.bound_evals, r.seconds_total);
    return 0;
}
