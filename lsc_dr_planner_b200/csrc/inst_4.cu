// inst_4.cu -- PDIP kernel instances, group 4 (see solve_instances.hpp)
#define LSCQP_TU 4
#include "solve_instances.hpp"
