// ceres/ceres.h — the slice of the Ceres 1.7.0 surface that SLSLAM's back end touches, re-pointed at the B200 solver.
//
// The reference compiles src/slam.cpp, src/lba_problem.cpp and src/po_problem.cpp against the real Ceres and calls
//     ceres::Problem problem;  xx_problem.build(&problem);
//     ceres::Solver::Options options;  xx_problem.set_options(&options);
//     ceres::Solver::Summary summary;  ceres::Solve(options, &problem, &summary);
// (reference src/slam.cpp:658-663, 939-944, 1288-1293).  With this header first on the include path the same
// caller code compiles unchanged; Problem only records which LBAProblem / POProblem was built into it and
// ceres::Solve forwards to the C ABI (include/slslam_b200.h).  Nothing here evaluates a cost function on the host:
// there is no CPU solver behind this shim.
#ifndef SLSLAM_B200_CERES_SHIM_H_
#define SLSLAM_B200_CERES_SHIM_H_

#include <cstdio>
#include <string>

#include "../slslam_b200.h"
#include "rotation.h"

namespace ceres {

enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum LoggingType { SILENT, PER_MINIMIZER_ITERATION };
// Solver::Summary::termination_type values this path can produce (Ceres 1.7.0 names) plus DID_NOT_RUN.
enum SolverTerminationType {
  NO_CONVERGENCE = SLSLAM_NO_CONVERGENCE,
  GRADIENT_TOLERANCE = SLSLAM_GRADIENT_TOLERANCE,
  FUNCTION_TOLERANCE = SLSLAM_FUNCTION_TOLERANCE,
  PARAMETER_TOLERANCE = SLSLAM_PARAMETER_TOLERANCE,
  NUMERICAL_FAILURE = SLSLAM_NUMERICAL_FAILURE,
  DID_NOT_RUN = 100
};

// LBAProblem::set_options fills one of these with every line and camera block in group 0
// (reference src/lba_problem.cpp:113-122).  The device solver always eliminates the line blocks first, which is the
// same step, so the ordering is accepted and ignored.
class ParameterBlockOrdering {
 public:
  ParameterBlockOrdering() : num_elements_(0) {}
  bool AddElementToGroup(const double*, int) { ++num_elements_; return true; }
  int NumElements() const { return num_elements_; }
 private:
  int num_elements_;
};

// What LBAProblem::build / POProblem::build leave behind instead of a list of residual blocks.
class Problem {
 public:
  enum Kind { EMPTY, LBA, PO };
  Problem() : kind(EMPTY), parameters(0) { lba = slslam_lba_desc(); po = slslam_po_desc(); }
  int NumResidualBlocks() const { return kind == LBA ? lba.num_observations : kind == PO ? po.num_edges : 0; }
  int NumParameters() const { return kind == LBA ? 6 * lba.num_cameras + 4 * lba.num_lines : kind == PO ? 6 * po.num_poses : 0; }
  Kind kind;
  slslam_lba_desc lba;
  slslam_po_desc po;
  double* parameters;   // the caller's parameter array, updated in place by Solve
};

class Solver {
 public:
  struct Options {
    Options()
        : linear_solver_type(SPARSE_NORMAL_CHOLESKY), num_linear_solver_threads(1), linear_solver_ordering(0),
          max_num_iterations(50), minimizer_progress_to_stdout(false), num_threads(1), eta(1e-1), logging_type(SILENT),
          function_tolerance(1e-6), gradient_tolerance(1e-10), parameter_tolerance(1e-8),
          initial_trust_region_radius(1e4) {}
    ~Options() { delete linear_solver_ordering; }
    LinearSolverType linear_solver_type;
    int num_linear_solver_threads;
    ParameterBlockOrdering* linear_solver_ordering;   // owned, as in Ceres 1.7.0
    int max_num_iterations;
    bool minimizer_progress_to_stdout;
    int num_threads;
    double eta;
    LoggingType logging_type;
    double function_tolerance, gradient_tolerance, parameter_tolerance, initial_trust_region_radius;
   private:
    Options(const Options&);
    Options& operator=(const Options&);
  };

  struct Summary {
    Summary()
        : termination_type(DID_NOT_RUN), initial_cost(-1.0), final_cost(-1.0), fixed_cost(-1.0),
          num_successful_steps(-1), num_unsuccessful_steps(-1), error_code(0) {}
    SolverTerminationType termination_type;
    double initial_cost, final_cost, fixed_cost;
    int num_successful_steps, num_unsuccessful_steps;
    int error_code;        // SLSLAM_OK or the negative code of the C ABI
    std::string error;
    std::string BriefReport() const {
      char buf[256];
      if (error_code != 0) snprintf(buf, sizeof(buf), "slslam_b200: solve failed (%d): %s", error_code, error.c_str());
      else snprintf(buf, sizeof(buf), "slslam_b200: iterations %d, initial cost %.6e, final cost %.6e, termination %d",
                    num_successful_steps + num_unsuccessful_steps, initial_cost, final_cost, (int)termination_type);
      return std::string(buf);
    }
    std::string FullReport() const { return BriefReport(); }
  };
};

// The reference ignores the outcome of Solve (void return, termination type never read: SURVEY.md §8b); on any
// failure the parameters are left untouched and Summary::error_code / error say why.  Never throws.
inline void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary) {
  Solver::Summary local;
  Solver::Summary* out = summary ? summary : &local;
  *out = Solver::Summary();
  if (!problem || problem->kind == Problem::EMPTY) return;
  slslam_summary s = slslam_summary();
  int rc;
  if (problem->kind == Problem::LBA) {
    slslam_lba_desc d = problem->lba;
    d.max_iterations = options.max_num_iterations;
    d.function_tolerance = options.function_tolerance; d.gradient_tolerance = options.gradient_tolerance;
    d.parameter_tolerance = options.parameter_tolerance; d.initial_trust_region_radius = options.initial_trust_region_radius;
    rc = slslam_lba_solve(&d, problem->parameters, &s);
  } else {
    slslam_po_desc d = problem->po;
    d.max_iterations = options.max_num_iterations;
    d.function_tolerance = options.function_tolerance; d.gradient_tolerance = options.gradient_tolerance;
    d.parameter_tolerance = options.parameter_tolerance; d.initial_trust_region_radius = options.initial_trust_region_radius;
    rc = slslam_po_solve(&d, problem->parameters, &s);
  }
  out->error_code = rc;
  if (rc != SLSLAM_OK) {
    out->error = std::string(slslam_strerror(rc)) + " [" + slslam_last_error() + "]";
    out->termination_type = NUMERICAL_FAILURE;
    // The reference adds these fields into its running statistics without looking at the outcome
    // (src/slam.cpp:949-952), so a failed solve contributes zeros, never the -1 "did not run" markers; and because
    // LBAProblem::set_options forces SILENT (lba_problem.cpp:130-131) a failure is reported on stderr regardless of it:
    // a BA that silently became a no-op (e.g. --ba_window_size beyond slslam_lba_get_limits) must be visible.
    out->initial_cost = out->final_cost = out->fixed_cost = 0.0;
    out->num_successful_steps = out->num_unsuccessful_steps = 0;
    fprintf(stderr, "%s\n", out->BriefReport().c_str());
    return;
  }
  out->termination_type = (SolverTerminationType)s.termination_type;
  out->initial_cost = s.initial_cost; out->final_cost = s.final_cost; out->fixed_cost = s.fixed_cost;
  out->num_successful_steps = s.num_successful_steps; out->num_unsuccessful_steps = s.num_unsuccessful_steps;
  if (options.minimizer_progress_to_stdout && options.logging_type != SILENT) printf("%s\n", out->BriefReport().c_str());
}

}  // namespace ceres

#endif  // SLSLAM_B200_CERES_SHIM_H_
