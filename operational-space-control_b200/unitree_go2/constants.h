#pragma once
// Derived sizes, same names as the reference's unitree_go2/constants.h:11-23.
#include "operational-space-control/unitree_go2/autogen/autogen_defines.h"

namespace operational_space_controller::constants {
    namespace optimization {
        constexpr int s_size = 6 * model::body_ids_size;   // stacked spatial task vector
        constexpr int p_size = 3 * model::body_ids_size;   // translational part
        constexpr int r_size = 3 * model::body_ids_size;   // rotational part
        constexpr int constraint_matrix_rows = Aeq_rows + Aineq_rows + design_vector_size;
        constexpr int constraint_matrix_cols = design_vector_size;
        constexpr int bounds_size = beq_sz + bineq_sz + design_vector_size;
    }
}
