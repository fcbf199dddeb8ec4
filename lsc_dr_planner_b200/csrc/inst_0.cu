// inst_0.cu -- PDIP kernel instances, group 0 (see solve_instances.hpp)
#define LSCQP_TU 0
#include "solve_instances.hpp"
