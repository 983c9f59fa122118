// sfm driver: same pipeline and defaults as the reference's main() (reference src/sfm.cc:77-131)
//   read -> hemisphere fit -> PLY -> solve(100 it, cameras frozen) -> filter ->
//   loop { solve(100 it); filter; PLY } until the point count is stable -> PLY + .deeparc
// The reference has no run-time flags (paths are compile-time macros, sfm.cc:22-27); with no
// arguments this binary behaves the same (same default paths), and accepts optional overrides:
//   sfm [--input F] [--output F] [--ply-init F] [--ply-adjust PREFIX] [--ply-clear F]
//       [--max-iter N] [--max-seconds S] [--filter-boundary E] [--max-outer N]
//       [--cauchy-loss A]      robust loss CauchyLoss(A) (commented out in the reference, sfm.cc:49)
//       [--output-binary F]    additionally write the lossless binary side format (read() accepts it as --input)
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "DeepArcManager.hh"
#include "ba_client.hh"

#define DEEPARC_INPUT "../data/teabottle_green_bfs.deeparc"
#define DEEPARC_OUTPUT "../assets/teabottle_green_bfs_.deeparc"
#define PLY_INIT "../assets/teabottle_green_bfs_init.ply"
#define PLY_CLEAR "../assets/teabottle_green_bfs_clear.ply"
#define PLY_ADJUST "../assets/teabottle_gree_bfs_adjust_point_"

int main(int argc, char** argv) {
  std::string input = DEEPARC_INPUT, output = DEEPARC_OUTPUT, ply_init = PLY_INIT, ply_clear = PLY_CLEAR,
              ply_adjust = PLY_ADJUST;
  int max_iter = 100, max_seconds = 3600, max_outer = 1 << 30;
  double boundary = 5.0;
  std::string output_binary;
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string k = argv[i], v = argv[i + 1];
    if (k == "--input") input = v;
    else if (k == "--output") output = v;
    else if (k == "--ply-init") ply_init = v;
    else if (k == "--ply-adjust") ply_adjust = v;
    else if (k == "--ply-clear") ply_clear = v;
    else if (k == "--max-iter") max_iter = std::atoi(v.c_str());
    else if (k == "--max-seconds") max_seconds = std::atoi(v.c_str());
    else if (k == "--filter-boundary") boundary = std::atof(v.c_str());
    else if (k == "--max-outer") max_outer = std::atoi(v.c_str());
    else if (k == "--cauchy-loss") solve_set_cauchy_loss(std::atof(v.c_str()));
    else if (k == "--output-binary") output_binary = v;
    else {
      std::cerr << "unknown option " << k << "\n";
      return 2;
    }
  }
  try {
    DeepArcManager deeparcManager;
    int old_point = 1, current_point = 10000000;  // sfm.cc:82
    deeparcManager.read(input);

    // hemisphere fit (sfm.cc:86-103): centre 0, rho 1, 1000 iterations
    std::vector<std::vector<double> > camera_center = deeparcManager.getCameraCenter();
    double hemisphere_center[3] = {0, 0, 0};
    double hemisphere_radius = 1;
    {
      std::vector<double> flat;
      for (const std::vector<double>& c : camera_center) flat.insert(flat.end(), c.begin(), c.end());
      dba_solve_options o;
      dba_solve_options_default(&o);
      o.max_num_iterations = 1000;
      o.max_solver_time_in_seconds = 3600;
      o.progress_to_stdout = 1;
      dba_summary s;
      std::memset(&s, 0, sizeof s);
      deeparc::check(dba_fit_hemisphere(deeparc::engine(), flat.data(), static_cast<int32_t>(camera_center.size()),
                                        hemisphere_center, &hemisphere_radius, &o, &s),
                     "dba_fit_hemisphere");
      std::cout << "hemisphere fit: " << s.message << " cost " << s.initial_cost << " -> " << s.final_cost << "\n";
    }
    std::cout << hemisphere_center[0] << " " << hemisphere_center[1] << " " << hemisphere_center[2] << " "
              << hemisphere_radius << "\n";

    deeparcManager.writePly(ply_init);
    solve(deeparcManager, max_iter, max_seconds, true);  // sfm.cc:111 points-only pass
    deeparcManager.filterPoint3d(boundary, hemisphere_center, hemisphere_radius);
    std::cout << "block: " << deeparcManager.parameters()->size() << "\n";
    std::cout << "point3d: " << deeparcManager.point3ds()->size() << "\n";
    int step = 0;
    deeparcManager.writePly(ply_adjust + std::to_string(step) + ".ply");
    while (current_point != old_point && step < max_outer) {  // sfm.cc:118-127
      step++;
      old_point = current_point;
      solve(deeparcManager, max_iter, max_seconds);
      deeparcManager.filterPoint3d(boundary, hemisphere_center, hemisphere_radius);
      std::cout << "block: " << deeparcManager.parameters()->size() << "\n";
      std::cout << "point3d: " << deeparcManager.point3ds()->size() << "\n";
      current_point = static_cast<int>(deeparcManager.point3ds()->size());
      deeparcManager.writePly(ply_adjust + std::to_string(step) + ".ply");
    }
    std::cout << "TOTAL REPEAT: " << step << "\n";
    deeparcManager.writePly(ply_clear);
    deeparcManager.write(output);
    if (!output_binary.empty()) deeparcManager.writeBinary(output_binary);
    deeparc::engine_release();
  } catch (const char* msg) {
    std::cerr << "sfm: " << msg << "\n";
    return 1;
  } catch (const std::exception& e) {
    std::cerr << "sfm: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
