// inst_1.cu -- PDIP kernel instances, group 1 (see solve_instances.hpp)
#define LSCQP_TU 1
#include "solve_instances.hpp"
