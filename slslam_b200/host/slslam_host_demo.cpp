// Host-side demo / test driver of the drop-in boundary: runs one LBA window or one pose graph through the SAME
// five-step call sequence the reference's SLAM class uses (reference src/slam.cpp:924-944 for LBA, :1283-1293 for PO):
//   construct -> setters -> build(&problem) -> set_options(&options) -> ceres::Solve(options, &problem, &summary)
// with include/lba_problem.h, include/po_problem.h and include/ceres/ceres.h standing where the reference's headers
// and the real Ceres stood.  tests/test_host_cpp.py feeds it seeded windows and checks the result against the oracle.
//
//   slslam_host_demo lba <in.bin> <out.bin>     in: int32 C L N max_iters robust | cam_idx[N] line_idx[N] fixed[2N] | obs[8N] params[6C+4L]
//   slslam_host_demo po  <in.bin> <out.bin>     in: int32 K E max_iters | idx1[E] idx2[E] | constraints[6E] params[6K]
//   out: double error_code initial_cost final_cost successful unsuccessful termination | params
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "lba_problem.h"
#include "po_problem.h"

static bool read_all(FILE* f, void* dst, size_t bytes) { return bytes == 0 || fread(dst, 1, bytes, f) == bytes; }

static int write_result(const char* path, const ceres::Solver::Summary& s, const double* params, int n) {
  FILE* f = fopen(path, "wb");
  if (!f) return 2;
  const double head[6] = {(double)s.error_code, s.initial_cost, s.final_cost, (double)s.num_successful_steps,
                          (double)s.num_unsuccessful_steps, (double)s.termination_type};
  fwrite(head, sizeof(double), 6, f);
  fwrite(params, sizeof(double), (size_t)n, f);
  fclose(f);
  return 0;
}

static int run_lba(FILE* in, const char* out_path) {
  int hdr[5];
  if (!read_all(in, hdr, sizeof(hdr))) return 2;
  const int num_cameras = hdr[0], num_lines = hdr[1], num_observations = hdr[2];
  const int num_parameters = 6 * num_cameras + 4 * num_lines;
  FLAGS_robust = hdr[4] != 0;
  // the caller allocates with new[]; LBAProblem takes ownership (reference src/lba_problem.cpp:46-52)
  int* camera_index = new int[num_observations];
  int* line_index = new int[num_observations];
  int* fixed_index = new int[2 * num_observations];
  double* observations = new double[8 * (size_t)num_observations];
  double* parameters = new double[num_parameters];
  if (!read_all(in, camera_index, 4 * (size_t)num_observations) || !read_all(in, line_index, 4 * (size_t)num_observations) ||
      !read_all(in, fixed_index, 8 * (size_t)num_observations) || !read_all(in, observations, 64 * (size_t)num_observations) ||
      !read_all(in, parameters, 8 * (size_t)num_parameters)) return 2;

  ceres::lba_param_t param;
  param.num_cameras = num_cameras;
  param.num_lines = num_lines;
  param.num_observations = num_observations;
  param.num_iterations = hdr[3];
  param.num_parameters = num_parameters;
  param.mode = MODE_SPARSE_SCHUR;

  ceres::LBAProblem ba_problem(param);
  ba_problem.set_line_index(line_index);
  ba_problem.set_camera_index(camera_index);
  ba_problem.set_fixed_index(fixed_index);
  ba_problem.set_observations(observations);
  ba_problem.set_parameters(parameters);

  ceres::Problem problem;
  ba_problem.build(&problem);
  ceres::Solver::Options options;
  ba_problem.set_options(&options);
  ceres::Solver::Summary summary;
  ceres::Solve(options, &problem, &summary);

  printf("%s\n", summary.BriefReport().c_str());
  // the caller keeps reading `parameters` until ba_problem goes out of scope (reference src/slam.cpp:957-972)
  return write_result(out_path, summary, parameters, num_parameters);
}

static int run_po(FILE* in, const char* out_path) {
  int hdr[3];
  if (!read_all(in, hdr, sizeof(hdr))) return 2;
  const int kfs_size = hdr[0], edge_size = hdr[1];
  int* pose_index_1 = new int[edge_size];
  int* pose_index_2 = new int[edge_size];
  double* constraints = new double[6 * (size_t)edge_size];
  double* parameters = new double[6 * (size_t)kfs_size];
  if (!read_all(in, pose_index_1, 4 * (size_t)edge_size) || !read_all(in, pose_index_2, 4 * (size_t)edge_size) ||
      !read_all(in, constraints, 48 * (size_t)edge_size) || !read_all(in, parameters, 48 * (size_t)kfs_size)) return 2;

  ceres::POProblem po_problem(edge_size, hdr[2]);
  po_problem.set_pose_index_1(pose_index_1);
  po_problem.set_pose_index_2(pose_index_2);
  po_problem.set_constraints(constraints);
  po_problem.set_parameters(parameters);
  ceres::Problem problem;
  po_problem.build(&problem);
  ceres::Solver::Options options;
  po_problem.set_options(&options);
  ceres::Solver::Summary summary;
  ceres::Solve(options, &problem, &summary);

  printf("%s\n", summary.BriefReport().c_str());
  return write_result(out_path, summary, parameters, 6 * kfs_size);
}

int main(int argc, char** argv) {
  if (argc != 4) { fprintf(stderr, "usage: %s lba|po <in.bin> <out.bin>\n", argv[0]); return 64; }
  FILE* in = fopen(argv[2], "rb");
  if (!in) { perror(argv[2]); return 2; }
  int rc = 64;
  if (!strcmp(argv[1], "lba")) rc = run_lba(in, argv[3]);
  else if (!strcmp(argv[1], "po")) rc = run_po(in, argv[3]);
  fclose(in);
  return rc;
}
