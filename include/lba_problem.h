// lba_problem.h — drop-in for the reference's src/lba_problem.h + src/lba_problem.cpp.
//
// Same public surface (lba_param_t, MODE_* macros, class ceres::LBAProblem with its accessors, five setters,
// set_logging_type, build(Problem*), set_options(Solver::Options*); reference src/lba_problem.h:32-33, 123-197), same
// ownership rule (the destructor delete[]s the five arrays the caller new[]ed; src/lba_problem.cpp:46-52) and same
// flag (FLAGS_robust read in the constructor; src/lba_problem.cpp:35).  What changes is what build() does: instead of
// one AutoDiffCostFunction + HuberLoss per observation (src/lba_problem.cpp:54-93) it records the arrays in the
// Problem, and ceres::Solve hands them to the device solver through slslam_lba_solve (include/slslam_b200.h).
// The LineReprojectionError functor itself lives on the device (slslam_b200/csrc/lba_math.cuh).
// Header-only: link the caller with -lslslam_b200.
#ifndef SLSLAM_B200_LBA_PROBLEM_H_
#define SLSLAM_B200_LBA_PROBLEM_H_

#include <string>

#include "ceres/ceres.h"

// same values as the reference (src/lba_problem.h:32-33)
#define MODE_SPARSE_SCHUR 1
#define MODE_SPARSE_NORMAL_CHOLESKY 2

// The reference declares the flag with gflags (DECLARE_bool(robust), defined in src/main.cpp:27).  With gflags on the
// include path that declaration is used as is; without it a plain global with the same name and default stands in.
#if defined(SLSLAM_B200_USE_GFLAGS)
#include <gflags/gflags.h>
DECLARE_bool(robust);
#else
#ifndef SLSLAM_B200_FLAGS_ROBUST_DEFINED
#define SLSLAM_B200_FLAGS_ROBUST_DEFINED
inline bool& slslam_b200_flags_robust() { static bool v = true; return v; }
#define FLAGS_robust (slslam_b200_flags_robust())
#endif
#endif

namespace ceres {

typedef struct {
  int num_cameras;
  int num_lines;
  int num_observations;
  int num_iterations;
  int num_parameters;   // 6 * num_cameras + 4 * num_lines
  int mode;             // accepted and ignored: the reference's switch falls through to one solver (lba_problem.cpp:96-101)
} lba_param_t;

class LBAProblem {
 public:
  explicit LBAProblem(lba_param_t param)
      : mode_(param.mode), num_cameras_(param.num_cameras), num_lines_(param.num_lines),
        num_observations_(param.num_observations), num_parameters_(param.num_parameters),
        num_iterations_(param.num_iterations), num_threads(1), eta(1e-2), robustify(FLAGS_robust), logging_type(false),
        line_index_(0), camera_index_(0), fixed_index_(0), observations_(0), parameters_(0) {}
  ~LBAProblem() {
    delete[] line_index_;
    delete[] camera_index_;
    delete[] fixed_index_;
    delete[] observations_;
    delete[] parameters_;
  }

  int camera_block_size() const { return 6; }
  int line_block_size() const { return 4; }
  int num_cameras() const { return num_cameras_; }
  int num_lines() const { return num_lines_; }
  int num_observations() const { return num_observations_; }
  int num_parameters() const { return num_parameters_; }
  const int* line_index() const { return line_index_; }
  const int* camera_index() const { return camera_index_; }
  const int* fixed_index() const { return fixed_index_; }
  const double* observations() const { return observations_; }
  const double* parameters() const { return parameters_; }
  double* mutable_cameras() { return parameters_; }
  double* mutable_lines() { return parameters_ + camera_block_size() * num_cameras_; }

  void set_line_index(int* idx) { line_index_ = idx; }
  void set_camera_index(int* idx) { camera_index_ = idx; }
  void set_fixed_index(int* idx) { fixed_index_ = idx; }
  void set_observations(double* d) { observations_ = d; }
  void set_parameters(double* d) { parameters_ = d; }
  void set_logging_type(bool b) { logging_type = b; }

  // Records the window in `problem`; the arrays stay owned by this object and must outlive the Solve call, exactly
  // as the residual blocks of the reference point into them.
  void build(Problem* problem) {
    problem->kind = Problem::LBA;
    slslam_lba_desc& d = problem->lba;
    d = slslam_lba_desc();
    d.num_cameras = num_cameras_; d.num_lines = num_lines_; d.num_observations = num_observations_;
    d.max_iterations = num_iterations_;
    d.camera_index = camera_index_; d.line_index = line_index_; d.fixed_index = fixed_index_;
    d.observations = observations_;
    d.robust = robustify ? 1 : 0;
    d.huber_delta = 1.0 / 406.05;   // HuberLoss(1.0 / focal_length), reference src/lba_problem.cpp:78-80
    d.baseline = 0.12;              // literal in the cost functor, reference src/lba_problem.h:101
    problem->parameters = parameters_;
  }

  void set_options(Solver::Options* options) {
    options->linear_solver_type = SPARSE_NORMAL_CHOLESKY;   // what the fall-through switch always selects
    options->num_linear_solver_threads = num_threads;
    delete options->linear_solver_ordering;
    options->linear_solver_ordering = new ParameterBlockOrdering;
    for (int i = 0; i < num_lines_; ++i) options->linear_solver_ordering->AddElementToGroup(mutable_lines() + 4 * i, 0);
    for (int i = 0; i < num_cameras_; ++i) options->linear_solver_ordering->AddElementToGroup(mutable_cameras() + 6 * i, 0);
    options->max_num_iterations = num_iterations_;
    options->minimizer_progress_to_stdout = true;
    options->num_threads = num_threads;
    options->eta = eta;
    if (logging_type == false) options->logging_type = SILENT;
  }

 private:
  LBAProblem(const LBAProblem&);
  LBAProblem& operator=(const LBAProblem&);

  int mode_, num_cameras_, num_lines_, num_observations_, num_parameters_, num_iterations_;
  int num_threads;
  double eta;
  bool robustify, logging_type;
  int* line_index_;
  int* camera_index_;
  int* fixed_index_;
  double* observations_;
  double* parameters_;   // [camera_0 .. camera_{C-1}, line_0 .. line_{L-1}]
};

}  // namespace ceres

#endif  // SLSLAM_B200_LBA_PROBLEM_H_
