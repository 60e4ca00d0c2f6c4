// asph_run — the reference's desktop command line, headless, as a native program over the C ABI of include/asph.h.
//
//   asph_run run SIMULATION_CONFIG SCENE_CONFIG [-s|--max-seconds S] [-c|--overwrite-config-file F] [-p|--statistics-enabled]
//                [-w|--statistics-path F] [--max-steps N] [--split-patterns F] [--dump F] [--vtk-dir D [--vtk-every N]]
//                [--gpus N]   N > 1: one process per GPU is started (this program again, once per rank), the fluid is cut
//                             into N x-slabs; rank 0 reports
//                [--restart-vtk SNAPSHOT] [--lib LIBRARY] [-q]
//   asph_run image JOB_FILE... [--out-dir D] [--only K]... [--max-steps N] [--split-patterns F] [--lib LIBRARY] [-q]
//
// mirrors the clap definition and the flow of platform/desktop/main_loop.rs:25-189, 209-358 for `run`: read the YAML files,
// optional key-wise overwrite, init_simulation_params, load ./split-patterns.yaml, init_fluid_sim, then single_step until
// the simulated time reaches --max-seconds.  Window, renderer and UI thread are out of scope.  The library
// is libasph_b200.so next to this program (../csrc/) — the CUDA path, which fails loudly without a GPU; --lib binds another
// library exporting the same ABI (the tests hand it the CPU oracle to check this host logic without a GPU).
// `--dump F` writes the final state: "ASPHDUMP", u64 n, then position[2n], velocity[2n], mass[n] as float32.  `--vtk-dir`
// writes the snapshots of the reference's VtkExporter (vtk.hpp).  `image` runs the reference's batch export jobs
// (ImageExportConfig records, platform/desktop/animation/mod.rs:28-288) with VTK snapshots in place of the Cairo / ffmpeg
// rendering: `<png_file>.vtk` for a still, `<png_file>.frames/file-000000.vtk ...` with positions interpolated between the
// two physics steps around each frame time for a video, `<png_file>.stat` when `output_stats` is set.
//
// Test hooks: `asph_run yaml-dump FILE` (the parse as JSON), `asph_run params-dump CONFIG [OVERWRITE]` (asph_params bytes,
// hex), `asph_run scene-dump SCENE FILE` (the particles add_fluid_block generates, same layout as --dump).
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "host.hpp"
#include "vtk.hpp"

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

void mkdirs(const std::string& path) {
  std::string cur;
  for (size_t i = 0; i <= path.size(); i++) {
    if (i == path.size() || path[i] == '/') {
      if (!cur.empty() && !exists(cur) && mkdir(cur.c_str(), 0777) != 0 && !exists(cur)) throw std::runtime_error("cannot create " + cur);
    }
    if (i < path.size()) cur += path[i];
  }
}
std::string dirname_of(const std::string& p) {
  const size_t s = p.find_last_of('/');
  return s == std::string::npos ? "." : (s == 0 ? "/" : p.substr(0, s));
}
std::string join(const std::string& dir, const std::string& p) { return (!p.empty() && p[0] == '/') ? p : dir + "/" + p; }
std::string json_str(const std::string& s) {
  std::string o = "\"";
  for (char c : s) { if (c == '"' || c == '\\') o += '\\'; o += c; }
  return o + "\"";
}

// One ImageExportConfig record, animation/mod.rs:75-288.  Returns whether the job reached its export time.
bool run_job(const yaml_lite::Node& job, const std::string& job_dir, const std::string& out_dir, host::Library& lib,
             const host::SplitPatterns& split, long max_steps, bool quiet) {
  for (const char* k : {"time", "config_path", "visualization_params", "png_file"})
    if (!job.find(k)) throw std::runtime_error(std::string("failed parsing export config file: missing field `") + k + "`");
  if (!job.at("visualization_params").find("visualized_attribute"))
    throw std::runtime_error("failed parsing export config file: missing field `visualized_attribute`");
  const std::string png_file = job.at("png_file").as_string();
  yaml_lite::Node cfg = yaml_lite::parse_file(join(job_dir, job.at("config_path").as_string()));
  const yaml_lite::Node* scene_inline = job.find("scene");
  const yaml_lite::Node* scene_file = job.find("scene_file");
  const bool has_inline = scene_inline && !scene_inline->is_null(), has_file = scene_file && !scene_file->is_null();
  if (has_inline == has_file) throw std::runtime_error(std::string("expected either 'scene' or 'scene_file'") + (has_inline ? ". Not both!" : ""));
  const host::SceneConfig scene = host::scene_from_yaml(has_inline ? *scene_inline : yaml_lite::parse_file(join(job_dir, scene_file->as_string())));
  if (const yaml_lite::Node* upd = job.find("update_attributes")) host::merge_overwrite(cfg, *upd);  // "not able to find attribute"
  asph_params params = host::params_from_yaml(cfg);
  host::init_simulation_params(params, scene);
  const asph_boundary boundary = host::scene_boundary(scene, params.init_boundary_handler);
  host::FluidSimulation sim(lib, params, host::scene_particles(scene), boundary, &split, true);
  host::StatisticsRecorder rec;

  const yaml_lite::Node* vs = job.find("video_start_time");
  const bool video = vs && !vs->is_null();
  auto opt_f = [&](const char* k, float dflt) { const yaml_lite::Node* v = job.find(k); return (v && !v->is_null()) ? float(v->as_double()) : dflt; };
  const float fps = opt_f("video_fps", 60.f), speed = opt_f("video_speed", 1.f), end = float(job.at("time").as_double());
  float next_export = video ? float(vs->as_double()) : end;
  const yaml_lite::Node* poe = job.find("panic_on_end");
  const bool panic_on_end = poe && !poe->is_null() && poe->as_bool();
  const yaml_lite::Node* ost = job.find("output_stats");
  const bool output_stats = ost && !ost->is_null() && ost->as_bool();
  const std::string out_path = join(out_dir, png_file);
  mkdirs(dirname_of(out_path));
  const std::string frames_dir = out_path + ".frames";
  if (video) mkdirs(frames_dir);
  std::vector<std::pair<std::string, std::string>> frames;
  long steps = 0;
  bool finished = false;
  uint64_t first_count = sim.num_fluid_particles();
  while (!finished) {
    if (max_steps >= 0 && steps >= max_steps) break;
    const float time_before = float(sim.time());
    std::vector<float> pos_before;
    if (video) pos_before = sim.field(ASPH_F_POSITION, 2);
    const float dt = sim.single_step_without_adaptivity(params);
    const asph_step_info info = sim.step_info();
    rec.record_step(info);
    if (steps == 0) first_count = info.n_particles_begin;
    steps++;
    const float now = float(sim.time());
    if (panic_on_end && now > end) throw std::runtime_error(">>>>>>>>>>>> REACHED END BEFORE EXPORT <<<<<<<<<<<<");
    while (next_export <= now) {
      if (video) {
        const float a = (next_export - time_before) / (now - time_before);
        if (a < 0.f) throw std::runtime_error("negative interpolation");
        std::vector<float> pos = sim.field(ASPH_F_POSITION, 2);
        const float b = 1.f - a;
        for (size_t k = 0; k < pos.size(); k++) {
          const float u = a * pos[k], v = b * pos_before[k];
          pos[k] = u + v;
        }
        char name[64], t[64];
        std::snprintf(name, sizeof(name), "file-%06zu.vtk", frames.size());
        host::write_vtk_file(frames_dir + "/" + name, sim, boundary, &pos);
        std::snprintf(t, sizeof(t), "%.9g", double(next_export));
        frames.push_back({name, t});
        const float inc = 1.f / fps * speed;
        next_export = next_export + inc;
        if (now > end) { finished = true; break; }
      } else {
        host::write_vtk_file(out_path + ".vtk", sim, boundary);
        finished = true;
        break;
      }
    }
    if (finished) break;
    sim.single_step_adaptivity(params, dt);
    if (!quiet) std::printf("  step %ld: t=%.5f dt=%.3e n=%llu\n", steps, double(now), dt, (unsigned long long)info.n_particles_end);
  }
  if (video) {
    FILE* fp = std::fopen((frames_dir + "/frames.vtk.series").c_str(), "w");
    if (!fp) throw std::runtime_error("cannot write into " + frames_dir);
    std::fputs("{\n\"file-series-version\": \"1.0\",\n\"files\": [", fp);
    for (size_t k = 0; k < frames.size(); k++)
      std::fprintf(fp, "%s\n{ \"name\": \"%s\", \"time\": %s }", k ? "," : "", frames[k].first.c_str(), frames[k].second.c_str());
    std::fputs("\n]\n}", fp);
    std::fclose(fp);
  }
  if (output_stats) {
    FILE* fp = std::fopen((out_path + ".stat").c_str(), "w");
    if (!fp) throw std::runtime_error("cannot write " + out_path + ".stat");
    std::fputs(rec.write_statistics(sim).c_str(), fp);
    std::fclose(fp);
  }
  FILE* fp = std::fopen((out_path + ".job.json").c_str(), "w");
  if (!fp) throw std::runtime_error("cannot write " + out_path + ".job.json");
  std::fprintf(fp, "{\"png_file\": %s, \"finished\": %s, \"steps\": %ld, \"simulated_time\": %.9g, \"export_time\": %.9g, \"video\": %s, "
                   "\"frames\": %zu, \"particles_first_step\": %llu, \"particles_end\": %llu, \"visualized_attribute\": %s, \"backend\": %s}\n",
               json_str(png_file).c_str(), finished ? "true" : "false", steps, sim.time(), double(end), video ? "true" : "false", frames.size(),
               (unsigned long long)first_count, (unsigned long long)sim.num_fluid_particles(),
               json_str(job.at("visualization_params").at("visualized_attribute").as_string()).c_str(), json_str(lib.backend_name()).c_str());
  std::fclose(fp);
  return finished;
}

int cmd_image(const std::vector<std::string>& a) {
  std::vector<std::string> files;
  std::vector<long> only;
  std::string out_dir, split_path, lib_path;
  long max_steps = -1;
  bool quiet = false;
  for (size_t i = 0; i < a.size(); i++) {
    const std::string& s = a[i];
    auto value = [&]() -> std::string {
      if (i + 1 >= a.size()) throw std::runtime_error("option " + s + " needs a value");
      return a[++i];
    };
    if (s == "--out-dir") out_dir = value();
    else if (s == "--only") only.push_back(std::atol(value().c_str()));
    else if (s == "--max-steps") max_steps = std::atol(value().c_str());
    else if (s == "--split-patterns") split_path = value();
    else if (s == "--lib") lib_path = value();
    else if (s == "-q" || s == "--quiet") quiet = true;
    else if (!s.empty() && s[0] == '-') throw std::runtime_error("unknown option " + s);
    else files.push_back(s);
  }
  if (files.empty()) { std::fprintf(stderr, "usage: asph_run image JOB_FILE... [--out-dir D] [--only K] [--max-steps N] [--lib LIBRARY] [-q]\n"); return 2; }
  if (split_path.empty()) split_path = exists("./split-patterns.yaml") ? "./split-patterns.yaml" : exe_dir() + "/../data/split-patterns.yaml";  // mod.rs:105
  host::SplitPatterns split;
  host::split_patterns_from_yaml(yaml_lite::parse_file(split_path), split);
  if (lib_path.empty()) lib_path = exe_dir() + "/../csrc/libasph_b200.so";
  host::Library lib(lib_path);
  long done = 0, reached = 0;
  for (auto& file : files) {
    char real[4096];
    const std::string path = realpath(file.c_str(), real) ? std::string(real) : file;
    const yaml_lite::Node jobs = yaml_lite::parse_file(path);
    if (jobs.kind != yaml_lite::Node::Seq) throw std::runtime_error("failed parsing export config file: expected a list of jobs");
    for (size_t k = 0; k < jobs.seq.size(); k++) {
      if (!only.empty() && std::find(only.begin(), only.end(), long(k)) == only.end()) continue;
      const yaml_lite::Node* png = jobs.seq[k].find("png_file");
      std::printf("%s[%zu] -> %s\n", path.substr(path.find_last_of('/') + 1).c_str(), k, png ? png->as_string().c_str() : "?");
      reached += run_job(jobs.seq[k], dirname_of(path), out_dir.empty() ? dirname_of(path) : out_dir, lib, split, max_steps, quiet) ? 1 : 0;
      done++;
    }
  }
  std::printf("%ld job(s), %ld reached their export time\n", done, reached);
  return 0;
}

int usage() {
  std::fprintf(stderr,
               "usage: asph_run run SIMULATION_CONFIG SCENE_CONFIG [-s SECONDS] [-c OVERWRITE.yaml] [-p] [-w STATS_FILE]\n"
               "                    [--max-steps N] [--split-patterns FILE] [--dump FILE] [--vtk-dir DIR [--vtk-every N]] [--restart-vtk FILE] [--lib LIBRARY] [-q]\n"
               "                    [--gpus N]\n"
               "       asph_run image JOB_FILE... [--out-dir DIR] [--only K] [--max-steps N] [--lib LIBRARY] [-q]\n");
  return 2;
}

void hex_to_bytes(const char* hex, uint8_t* out, size_t n) {
  if (!hex || std::strlen(hex) != 2 * n) throw std::runtime_error("ASPH_RUN_NCCL_ID is missing or malformed");
  for (size_t k = 0; k < n; k++) {
    unsigned v = 0;
    std::sscanf(hex + 2 * k, "%2x", &v);
    out[k] = uint8_t(v);
  }
}

std::vector<std::string> g_argv;  // the command line, for the ranks of a multi-GPU run

// one child per GPU: the same command with ASPH_RUN_RANK / ASPH_RUN_WORLD / ASPH_RUN_NCCL_ID (and LOCAL_RANK for the device)
int spawn_ranks(int world, const uint8_t id[128]) {
  std::string hex(256, '0');
  for (int k = 0; k < 128; k++) std::snprintf(&hex[2 * size_t(k)], 3, "%02x", id[k]);
  hex.resize(256);
  std::vector<pid_t> pids;
  for (int r = 0; r < world; r++) {
    const pid_t pid = fork();
    if (pid < 0) throw std::runtime_error("fork failed");
    if (pid == 0) {
      setenv("ASPH_RUN_RANK", std::to_string(r).c_str(), 1);
      setenv("ASPH_RUN_WORLD", std::to_string(world).c_str(), 1);
      setenv("ASPH_RUN_NCCL_ID", hex.c_str(), 1);
      setenv("LOCAL_RANK", std::to_string(r).c_str(), 1);
      std::vector<char*> argv;
      for (auto& s : g_argv) argv.push_back(const_cast<char*>(s.c_str()));
      argv.push_back(nullptr);
      execv("/proc/self/exe", argv.data());
      std::perror("execv");
      _exit(127);
    }
    pids.push_back(pid);
  }
  int worst = 0;
  for (pid_t pid : pids) {
    int st = 0;
    waitpid(pid, &st, 0);
    const int rc = WIFEXITED(st) ? WEXITSTATUS(st) : 128;
    worst = std::max(worst, rc);
  }
  return worst;
}

int cmd_run(const std::vector<std::string>& a) {
  std::vector<std::string> positional;
  double max_seconds = -1;
  long max_steps = -1;
  std::string overwrite, stats_path, split_path, dump, lib_path, vtk_dir, restart_vtk;
  long vtk_every = 1;
  int gpus = 1;
  bool stats = false, quiet = false;
  for (size_t i = 0; i < a.size(); i++) {
    const std::string& s = a[i];
    auto value = [&]() -> std::string {
      if (i + 1 >= a.size()) throw std::runtime_error("option " + s + " needs a value");
      return a[++i];
    };
    if (s == "-s" || s == "--max-seconds") max_seconds = std::atof(value().c_str());
    else if (s == "--gpus") gpus = std::max(1, std::atoi(value().c_str()));
    else if (s == "-c" || s == "--overwrite-config-file") overwrite = value();
    else if (s == "-p" || s == "--statistics-enabled") stats = true;
    else if (s == "-w" || s == "--statistics-path") { stats_path = value(); stats = true; }
    else if (s == "--max-steps") max_steps = std::atol(value().c_str());
    else if (s == "--split-patterns") split_path = value();
    else if (s == "--dump") dump = value();
    else if (s == "--vtk-dir") vtk_dir = value();
    else if (s == "--vtk-every") vtk_every = std::max(1L, std::atol(value().c_str()));
    else if (s == "--restart-vtk") restart_vtk = value();
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
  // ---- several GPUs: the parent hands every rank the NCCL id in its environment and waits; a rank is this same command
  const char* rank_env = std::getenv("ASPH_RUN_RANK");
  const int rank = rank_env ? std::atoi(rank_env) : -1;
  if (gpus > 1 && rank < 0) {
    if (!dump.empty() || !vtk_dir.empty() || !restart_vtk.empty()) throw std::runtime_error("--dump / --vtk-dir / --restart-vtk need --gpus 1");
    uint8_t id[128];
    const int rc = lib.comm_unique_id(id);
    if (rc != ASPH_OK) throw std::runtime_error(std::string("asph_comm_unique_id failed: ") + host::status_name(rc) + " (NCCL not loadable?)");
    return spawn_ranks(gpus, id);
  }
  host::Particles particles;
  if (restart_vtk.empty()) {
    particles = host::scene_particles(scene);
  } else {  // restart: the particles come from a snapshot (x, v, m = the whole persistent state); the scene gives the boundary
    const host::VtkSnapshot snap = host::read_vtk_file(restart_vtk);
    particles.pos = snap.position; particles.vel = snap.get("velocity"); particles.mass = snap.get("mass");
    if (particles.vel.size() != particles.pos.size() || 2 * particles.mass.size() != particles.pos.size())
      throw std::runtime_error("snapshot arrays have inconsistent lengths: " + restart_vtk);
  }
  const asph_boundary boundary = host::scene_boundary(scene, params.init_boundary_handler);
  std::unique_ptr<host::FluidSimulation> sim_holder;
  if (rank >= 0) {  // this rank's contiguous share of the reference order (x-major lattices: already x-slabs)
    const int world = std::atoi(std::getenv("ASPH_RUN_WORLD"));
    uint8_t id[128];
    hex_to_bytes(std::getenv("ASPH_RUN_NCCL_ID"), id, 128);
    const uint64_t n_global = particles.n(), lo = n_global * uint64_t(rank) / uint64_t(world), hi = n_global * uint64_t(rank + 1) / uint64_t(world);
    host::Particles mine;
    mine.pos.assign(particles.pos.begin() + 2 * lo, particles.pos.begin() + 2 * hi);
    mine.vel.assign(particles.vel.begin() + 2 * lo, particles.vel.begin() + 2 * hi);
    mine.mass.assign(particles.mass.begin() + lo, particles.mass.begin() + hi);
    std::vector<uint32_t> gidx(size_t(hi - lo));
    for (uint64_t k = lo; k < hi; k++) gidx[size_t(k - lo)] = uint32_t(k);
    sim_holder.reset(new host::FluidSimulation(lib, params, mine, gidx, n_global, boundary, &split, stats, id, rank, world, rank));
    quiet = quiet || rank != 0;
  } else {
    sim_holder.reset(new host::FluidSimulation(lib, params, particles, boundary, &split, stats));  // init_fluid_sim, simulation.rs:3074
  }
  host::FluidSimulation& sim = *sim_holder;

  host::StatisticsRecorder rec;
  std::unique_ptr<host::VtkExporter> vtk;
  if (!vtk_dir.empty()) { mkdirs(vtk_dir); vtk.reset(new host::VtkExporter(vtk_dir, "my-sph")); }  // main_loop.rs:256
  long step = 0;
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    if (max_seconds >= 0 && sim.time() >= max_seconds) break;
    if (max_steps >= 0 && step >= max_steps) break;
    const auto ts = std::chrono::steady_clock::now();
    float dt;
    if (vtk && step % vtk_every == 0) {
      // the split form of the step, as the reference's exporter uses it: the per-step fields of a snapshot describe the
      // particle set of the physics step, before resampling changes it
      dt = sim.single_step_without_adaptivity(params);
      vtk->add_snapshot(sim.time(), sim, boundary);
      sim.single_step_adaptivity(params, dt);
    } else {
      dt = sim.single_step(params);  // params by value every step, main_loop.rs:280
    }
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
  if (rank >= 0) std::printf("[rank %d] ", rank);
  std::printf("%ld steps, simulated %.4f s, %llu particles%s, wall %.2f s, backend %s\n", step, sim.time(),
              (unsigned long long)sim.num_fluid_particles(), rank >= 0 ? " owned" : "", wall, lib.backend_name());
  if (stats && rank <= 0) {
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
  g_argv.assign(argv, argv + argc);
  std::vector<std::string> args(argv + 1, argv + argc);
  try {
    if (args.empty()) return usage();
    const std::string cmd = args[0];
    args.erase(args.begin());
    if (cmd == "run") return cmd_run(args);
    if (cmd == "image") return cmd_image(args);
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
    if (cmd == "generate-split-patterns") {
      std::fprintf(stderr, "`generate-split-patterns` (the offline pattern optimiser) is out of scope\n");
      return 2;
    }
    return usage();
  } catch (const std::exception& e) {
    std::fprintf(stderr, "asph_run: %s\n", e.what());
    return 1;
  }
}
