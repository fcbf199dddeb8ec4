// inst_3.cu -- PDIP kernel instances, group 3 (see solve_instances.hpp)
#define LSCQP_TU 3
#include "solve_instances.hpp"
