// asph_run — the reference's desktop command line, headless, as a native program over the C ABI of include/asph.h.
//
//   asph_run run SIMULATION_CONFIG SCENE_CONFIG [-s|--max-seconds S] [-c|--overwrite-config-file F] [-p|--statistics-enabled]
//                [-w|--statistics-path F] [--max-steps N] [--split-patterns F] [--dump F] [--lib LIBRARY] [-q]
//
// mirrors the clap definition and the flow of platform/desktop/main_loop.rs:25-189, 209-358 for `run`: read the YAML files,
// optional key-wise overwrite, init_simulation_params, load ./split-patterns.yaml, init_fluid_sim, then single_step until
// the simulated time reaches --max-seconds.  Window, renderer and UI thread are out of scope.  The library
// is libasph_b200.so next to this program (../csrc/) — the CUDA path, which fails loudly without a GPU; --lib binds another
// library exporting the same ABI (the tests hand it the CPU oracle to check this host logic without a GPU).
// `--dump F` writes the final state: "ASPHDUMP", u64 n, then position[2n], velocity[2n], mass[n] as float32.
//
// Test hooks: `asph_run yaml-dump FILE` (the parse as JSON), `asph_run params-dump CONFIG [OVERWRITE]` (asph_params bytes,
// hex), `asph_run scene-dump SCENE FILE` (the particles add_fluid_block generates, same layout as --dump).
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "host.hpp"

namespace {

std::string exe_dir() {
  char buf[4096];
  const ssize_t n = readlink("/proc/self/exe", buf, sizeof(buf) - 1);
  if (n <= 0) return ".";
  buf[n] = 0;
  std::string p(buf);
  const size_t s = p.find_last_of('/');
  return s == std::string::npos ? "." : p.substr(0, s);
}
bool exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0; }

void write_dump(const std::string& path, const std::vector<float>& pos, const std::vector<float>& vel, const std::vector<float>& mass) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("cannot write " + path);
  const uint64_t n = mass.size();
  std::fwrite("ASPHDUMP", 1, 8, f);
  std::fwrite(&n, sizeof(n), 1, f);
  std::fwrite(pos.data(), sizeof(float), pos.size(), f);
  std::fwrite(vel.data(), sizeof(float), vel.size(), f);
  std::fwrite(mass.data(), sizeof(float), mass.size(), f);
  std::fclose(f);
}

int usage() {
  std::fprintf(stderr,
               "usage: asph_run run SIMULATION_CONFIG SCENE_CONFIG [-s SECONDS] [-c OVERWRITE.yaml] [-p] [-w STATS_FILE]\n"
               "                    [--max-steps N] [--split-patterns FILE] [--dump FILE] [--lib LIBRARY] [-q]\n");
  return 2;
}

int cmd_run(const std::vector<std::string>& a) {
  std::vector<std::string> positional;
  double max_seconds = -1;
  long max_steps = -1;
  std::string overwrite, stats_path, split_path, dump, lib_path;
  bool stats = false, quiet = false;
  for (size_t i = 0; i < a.size(); i++) {
    const std::string& s = a[i];
    auto value = [&]() -> std::string {
      if (i + 1 >= a.size()) throw std::runtime_error("option " + s + " needs a value");
      return a[++i];
    };
    if (s == "-s" || s == "--max-seconds") max_seconds = std::atof(value().c_str());
    else if (s == "-c" || s == "--overwrite-config-file") overwrite = value();
    else if (s == "-p" || s == "--statistics-enabled") stats = true;
    else if (s == "-w" || s == "--statistics-path") { stats_path = value(); stats = true; }
    else if (s == "--max-steps") max_steps = std::atol(value().c_str());
    else if (s == "--split-patterns") split_path = value();
    else if (s == "--dump") dump = value();
    else if (s == "--lib") lib_path = value();
    else if (s == "-q" || s == "--quiet") quiet = true;
    else if (!s.empty() && s[0] == '-') throw std::runtime_error("unknown option " + s);
    else positional.push_back(s);
  }
  if (positional.size() != 2) return usage();
  if (max_seconds < 0 && max_steps < 0) { std::fprintf(stderr, "headless run needs --max-seconds or --max-steps\n"); return 2; }

  yaml_lite::Node cfg = yaml_lite::parse_file(positional[0]);
  if (!overwrite.empty()) host::merge_overwrite(cfg, yaml_lite::parse_file(overwrite));  // main_loop.rs:113-126
  asph_params params = host::params_from_yaml(cfg);
  const host::SceneConfig scene = host::scene_from_yaml(yaml_lite::parse_file(positional[1]));
  host::init_simulation_params(params, scene);
  if (split_path.empty()) split_path = exists("./split-patterns.yaml") ? "./split-patterns.yaml" : exe_dir() + "/../data/split-patterns.yaml";  // main_loop.rs:225
  host::SplitPatterns split;
  host::split_patterns_from_yaml(yaml_lite::parse_file(split_path), split);
  if (lib_path.empty()) lib_path = exe_dir() + "/../csrc/libasph_b200.so";
  host::Library lib(lib_path);
  const host::Particles particles = host::scene_particles(scene);
  const asph_boundary boundary = host::scene_boundary(scene, params.init_boundary_handler);
  host::FluidSimulation sim(lib, params, particles, boundary, &split, stats);  // init_fluid_sim, simulation.rs:3074

  host::StatisticsRecorder rec;
  long step = 0;
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    if (max_seconds >= 0 && sim.time() >= max_seconds) break;
    if (max_steps >= 0 && step >= max_steps) break;
    const auto ts = std::chrono::steady_clock::now();
    const float dt = sim.single_step(params);  // params by value every step, main_loop.rs:280
    const asph_step_info info = sim.step_info();
    rec.record_step(info);
    step++;
    if (!quiet) {
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ts).count();
      std::printf("step %ld: t=%.5f dt=%.3e n=%llu div-iters=%d density-iters=%d shared=%d merged=%d split=%d  %.2fms\n", step, sim.time(), dt,
                  (unsigned long long)info.n_particles_end, info.div_iterations, info.density_iterations, info.n_shared, info.n_merged,
                  info.n_split_parents, ms);
    }
  }
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf("%ld steps, simulated %.4f s, %llu particles, wall %.2f s, backend %s\n", step, sim.time(),
              (unsigned long long)sim.num_fluid_particles(), wall, lib.backend_name());
  if (stats) {
    const std::string text = rec.write_statistics(sim);
    if (!stats_path.empty()) {
      FILE* f = std::fopen(stats_path.c_str(), "w");
      if (!f) throw std::runtime_error("cannot write " + stats_path);
      std::fputs(text.c_str(), f);
      std::fclose(f);
    } else {
      std::fputs(text.c_str(), stdout);
    }
  }
  if (!dump.empty()) write_dump(dump, sim.field(ASPH_F_POSITION, 2), sim.field(ASPH_F_VELOCITY, 2), sim.field(ASPH_F_MASS, 1));
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  std::vector<std::string> args(argv + 1, argv + argc);
  try {
    if (args.empty()) return usage();
    const std::string cmd = args[0];
    args.erase(args.begin());
    if (cmd == "run") return cmd_run(args);
    if (cmd == "yaml-dump" && args.size() == 1) {
      std::string out;
      yaml_lite::to_json(yaml_lite::parse_file(args[0]), out);
      std::puts(out.c_str());
      return 0;
    }
    if (cmd == "params-dump" && (args.size() == 1 || args.size() == 2)) {
      yaml_lite::Node cfg = yaml_lite::parse_file(args[0]);
      if (args.size() == 2) host::merge_overwrite(cfg, yaml_lite::parse_file(args[1]));
      const asph_params p = host::params_from_yaml(cfg);
      const unsigned char* b = reinterpret_cast<const unsigned char*>(&p);
      for (size_t k = 0; k < sizeof(p); k++) std::printf("%02x", b[k]);
      std::printf("\n");
      return 0;
    }
    if (cmd == "scene-dump" && args.size() == 2) {
      const host::Particles p = host::scene_particles(host::scene_from_yaml(yaml_lite::parse_file(args[0])));
      write_dump(args[1], p.pos, p.vel, p.mass);
      return 0;
    }
    if (cmd == "image" || cmd == "generate-split-patterns") {
      std::fprintf(stderr, "`%s` is not part of the native host (batch export jobs: python asph_b200.py image; the pattern optimiser is out of scope)\n", cmd.c_str());
      return 2;
    }
    return usage();
  } catch (const std::exception& e) {
    std::fprintf(stderr, "asph_run: %s\n", e.what());
    return 1;
  }
}
