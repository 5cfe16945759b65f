// host_io.cpp -- placeholder translation unit for goicp_run_toml (filled in with the loaders).
#include "../../include/goicp_b200.h"
extern "C" int goicp_run_toml(const char*, unsigned, unsigned, goicp_result*) { return GOICP_ERR_IO; }
