// inst_2.cu -- PDIP kernel instances, group 2 (see solve_instances.hpp)
#define LSCQP_TU 2
#include "solve_instances.hpp"
