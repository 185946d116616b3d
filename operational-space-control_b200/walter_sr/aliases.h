#pragma once
// Same alias names as the reference's walter_sr/aliases.h:12-23 (row-major doubles).
#include "operational-space-control/compat/compat.h"
#include "operational-space-control/walter_sr/constants.h"

namespace operational_space_controller {
    namespace aliases {
        using namespace operational_space_controller::constants;
#if OSC_B200_HAVE_EIGEN
        template <int Rows_, int Cols_>
        using Matrix = Eigen::Matrix<double, Rows_, Cols_, Eigen::RowMajor>;
        template <int Rows_>
        using Vector = Eigen::Matrix<double, Rows_, 1>;
        template <int Rows_, int Cols_>
        using MatrixColMajor = Eigen::Matrix<double, Rows_, Cols_, Eigen::ColMajor>;
#else
        template <int Rows_, int Cols_>
        using Matrix = osc_b200::FixedMatrix<Rows_, Cols_, true>;
        template <int Rows_>
        using Vector = osc_b200::FixedMatrix<Rows_, 1, false>;
        template <int Rows_, int Cols_>
        using MatrixColMajor = osc_b200::FixedMatrix<Rows_, Cols_, false>;
#endif
        using TaskspaceTargets = Matrix<model::site_ids_size, 6>;
        using OptimizationSolution = Vector<optimization::design_vector_size>;
    }
}
