// po_problem.h — drop-in for the reference's src/po_problem.h + src/po_problem.cpp.
//
// Same public surface (class ceres::POProblem(int size, int iterations), accessors, setters, build, set_options;
// reference src/po_problem.h:110-144) and ownership (the destructor delete[]s the four arrays; src/po_problem.cpp:33-38).
// build() records the graph in the Problem instead of creating one AutoDiffCostFunction per edge
// (src/po_problem.cpp:40-65); ceres::Solve runs it on the device through slslam_po_solve.  The PoseConstraintError
// functor and its SE(3) helpers (src/po_problem.h:27-105) live on the device (slslam_b200/csrc/po_math.cuh).
// Header-only: link the caller with -lslslam_b200.
#ifndef SLSLAM_B200_PO_PROBLEM_H_
#define SLSLAM_B200_PO_PROBLEM_H_

#include "ceres/ceres.h"

namespace ceres {

class POProblem {
 public:
  explicit POProblem(int s, int n)
      : num_iterations(n), num_threads(1), eta(1e-2), robustify(false), size_(s), pose_index_1_(0), pose_index_2_(0),
        constraints_(0), parameters_(0) {}
  ~POProblem() {
    delete[] pose_index_1_;
    delete[] pose_index_2_;
    delete[] constraints_;
    delete[] parameters_;
  }

  int pose_block_size() const { return 6; }
  int num_size() const { return size_; }
  const int* pose_index_1() const { return pose_index_1_; }
  const int* pose_index_2() const { return pose_index_2_; }
  const double* constraints() const { return constraints_; }
  double* parameters() const { return parameters_; }

  void set_size(int s) { size_ = s; }
  void set_num_iterations(int s) { num_iterations = s; }
  void set_pose_index_1(int* idx) { pose_index_1_ = idx; }
  void set_pose_index_2(int* idx) { pose_index_2_ = idx; }
  void set_constraints(double* d) { constraints_ = d; }
  void set_parameters(double* d) { parameters_ = d; }

  // The reference never tells POProblem how many poses the parameter array holds (keyframe ids index it directly,
  // src/slam.cpp:1276-1280), and Ceres only ever touches the blocks an edge names; so the pose count handed to the
  // device is 1 + the largest index in the edge list, and poses beyond it are not read or written.
  void build(Problem* problem) {
    problem->kind = Problem::PO;
    slslam_po_desc& d = problem->po;
    d = slslam_po_desc();
    int K = 0;
    for (int i = 0; i < size_; ++i) {
      if (pose_index_1_[i] + 1 > K) K = pose_index_1_[i] + 1;
      if (pose_index_2_[i] + 1 > K) K = pose_index_2_[i] + 1;
    }
    d.num_poses = K; d.num_edges = size_; d.max_iterations = num_iterations;
    d.pose_index_1 = pose_index_1_; d.pose_index_2 = pose_index_2_; d.constraints = constraints_;
    problem->parameters = parameters_;
  }

  void set_options(Solver::Options* options) {
    options->linear_solver_type = SPARSE_NORMAL_CHOLESKY;
    options->num_linear_solver_threads = num_threads;
    options->max_num_iterations = num_iterations;
    options->minimizer_progress_to_stdout = true;
    options->num_threads = num_threads;
    options->eta = eta;
    options->logging_type = SILENT;
  }

 private:
  POProblem(const POProblem&);
  POProblem& operator=(const POProblem&);

  int num_iterations, num_threads;
  double eta;
  bool robustify;   // false: no loss function on pose-graph edges (src/po_problem.cpp:27,55)
  int size_;
  int* pose_index_1_;
  int* pose_index_2_;
  double* constraints_;
  double* parameters_;
};

}  // namespace ceres

#endif  // SLSLAM_B200_PO_PROBLEM_H_
