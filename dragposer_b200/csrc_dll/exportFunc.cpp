// exportFunc.cpp -- native DragPoserDLL C ABI over the B200 engine (see include/exportFunc.h).
#include "../../include/exportFunc.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/dp_engine.h"

namespace {

constexpr int J = DP_JOINTS;
constexpr size_t kTemporalFloats = 1283976;  // dp_engine_temporal_blob_floats(), checked at load time

struct Model {
  std::vector<float> A0, b0, A1, b1, A2, b2, mean_q, std_q, mean_d, std_d;
  std::vector<float> encA[3], encb[3], mu_w, mu_b, lv_w, lv_b, temporal, means_latent, stds_latent;
};

void log_line(const std::string& msg) {
  fprintf(stderr, "[DragPoserDLL] %s\n", msg.c_str());
  if (const char* path = getenv("DRAGPOSER_LOG")) {
    std::ofstream f(path, std::ios::app);
    f << msg << "\n";
  }
}

}  // namespace

struct DragPoser {
  dp_engine* engine = nullptr;
  int num_joints = 0, num_ee = 0;
  std::vector<int32_t> parents;
  std::vector<float> offsets;
  Model model;
  bool have_skeleton = false, have_model = false, have_session = false, have_latent_override = false;
  std::vector<int32_t> ee_joints;
  std::vector<float> ee_weights;
  float latent0[DP_LATENT] = {0};
  dp_run_params params{};
  int status = 0;
  std::string message;
  std::mt19937 rng{2222};  // train.param["seed"]; the reference draws eps from torch's RNG (run_drag.py:17, autoencoder.py:19-22)
  std::vector<float> tp, tr, pose;

  DragPoser() {
    params.stop_eps_pos = 1e-2;  // DragPose.run defaults (drag_pose.py:203-213) until set_optim_params/set_lambdas are called
    params.stop_eps_rot = 1e-2;
    params.min_loss_incr = 0.00001;
    params.max_iter = 100;
    params.learning_rate = 1e-3f;
    params.lambda_rot = 1.0f;
    params.lambda_temporal = 1.0f;
    params.temporal_future_window = 60;
    params.joint_adjust_joint = -1;  // run_drag.py:155 passes joint_adjustment_indices=None
    params.joint_adjust_slot = 0;
    params.joint_adjust_weight = 0.01f;
    params.decoder_path = 0;
  }
  int fail(const std::string& m) {
    status = -1;
    message = m;
    log_line(m);
    return -1;
  }
  void ok() {
    status = 0;
    message.clear();
  }
};

namespace {

bool parse_bvh_hierarchy(const std::string& path, std::vector<int32_t>& parents, std::vector<float>& offsets, std::string& err) {
  std::ifstream f(path);
  if (!f) {
    err = "cannot open BVH file " + path;
    return false;
  }
  std::vector<int> stack;
  std::string tok;
  bool in_end = false;
  int cur = -1;
  while (f >> tok) {
    if (tok == "MOTION") break;
    if (tok == "ROOT" || tok == "JOINT") {
      std::string name;
      f >> name;
      parents.push_back(stack.empty() ? 0 : stack.back());
      offsets.insert(offsets.end(), {0.f, 0.f, 0.f});
      cur = (int)parents.size() - 1;
    } else if (tok == "End") {
      f >> tok;  // "Site"
      in_end = true;
    } else if (tok == "{") {
      stack.push_back(in_end ? -1 : cur);
    } else if (tok == "}") {
      if (stack.empty()) {
        err = "unbalanced braces in BVH hierarchy";
        return false;
      }
      if (stack.back() == -1) in_end = false;
      stack.pop_back();
      cur = stack.empty() ? -1 : stack.back();
    } else if (tok == "OFFSET") {
      float x, y, z;
      f >> x >> y >> z;
      if (!in_end && cur >= 0) {
        offsets[3 * cur] = x;
        offsets[3 * cur + 1] = y;
        offsets[3 * cur + 2] = z;
      }
    }
  }
  if (parents.empty()) {
    err = "no joints found in " + path;
    return false;
  }
  parents[0] = 0;                                // train.py:338
  offsets[0] = offsets[1] = offsets[2] = 0.f;    // train.py:340
  return true;
}

bool read_dpm(const std::string& path, Model& m, bool& trained_temporal, std::string& err) {
  std::ifstream f(path, std::ios::binary);
  if (!f) {
    err = "cannot open " + path;
    return false;
  }
  char magic[4];
  uint32_t version = 0, n = 0;
  f.read(magic, 4);
  f.read(reinterpret_cast<char*>(&version), 4);
  f.read(reinterpret_cast<char*>(&n), 4);
  if (!f || memcmp(magic, "DPM1", 4) != 0 || (version != 1 && version != 2)) {
    err = path + " is not a DPM1 model file";
    return false;
  }
  uint32_t flags = 0;  // version 2: bit 0 = the predictor was read from a temporal.pt (export_model.py)
  if (version >= 2) f.read(reinterpret_cast<char*>(&flags), 4);
  trained_temporal = (flags & 1u) != 0;
  std::vector<float> flat(n);
  f.read(reinterpret_cast<char*>(flat.data()), (std::streamsize)n * 4);
  if (!f) {
    err = path + " is truncated";
    return false;
  }
  size_t o = 0;
  auto take = [&](std::vector<float>& dst, size_t cnt) {
    if (o + cnt > flat.size()) return false;
    dst.assign(flat.begin() + o, flat.begin() + o + cnt);
    o += cnt;
    return true;
  };
  const size_t enc_dims[4] = {176, 112, 72, 48};
  bool good = take(m.A0, 40 * 24) && take(m.b0, 40) && take(m.A1, 60 * 40) && take(m.b1, 60) && take(m.A2, 92 * 60) &&
              take(m.b2, 92) && take(m.mean_q, 88) && take(m.std_q, 88) && take(m.mean_d, 3) && take(m.std_d, 3);
  for (int l = 0; l < 3 && good; ++l) good = take(m.encA[l], enc_dims[l + 1] * enc_dims[l]) && take(m.encb[l], enc_dims[l + 1]);
  good = good && take(m.mu_w, 24 * 48) && take(m.mu_b, 24) && take(m.lv_w, 24 * 48) && take(m.lv_b, 24) &&
         take(m.temporal, dp_engine_temporal_blob_floats()) && take(m.means_latent, 24) && take(m.stds_latent, 24);
  if (!good || o != flat.size()) {
    err = path + " has an unexpected size";
    return false;
  }
  return true;
}

// Folded encoder on the host (3 masked linears + LeakyReLU, then mu / logvar heads); once per session.
void encode_zero_pose(const Model& m, float mu[24], float logvar[24]) {
  const size_t dims[4] = {176, 112, 72, 48};
  std::vector<float> x(176, 0.0f), y;  // RunDrag encodes the zero (= mean) standardised pose (run_drag.py:90)
  for (int l = 0; l < 3; ++l) {
    y.assign(dims[l + 1], 0.f);
    for (size_t o = 0; o < dims[l + 1]; ++o) {
      float a = m.encb[l][o];
      for (size_t i = 0; i < dims[l]; ++i) a += m.encA[l][o * dims[l] + i] * x[i];
      y[o] = a > 0.f ? a : 0.2f * a;
    }
    x = y;
  }
  for (int o = 0; o < 24; ++o) {
    float a = m.mu_b[o], b = m.lv_b[o];
    for (int i = 0; i < 48; ++i) {
      a += m.mu_w[o * 48 + i] * x[i];
      b += m.lv_w[o * 48 + i] * x[i];
    }
    mu[o] = a;
    logvar[o] = b;
  }
}

void quat_mul(const float* a, const float* b, float* r) {
  r[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  r[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  r[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  r[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}

void quat_to_matrix(const quaternion& q, float* m) {  // pymotion quat.to_matrix == utils.py:34-76 3x3 block
  const float w = q.w, x = q.x, y = q.y, z = q.z, x2 = x + x, y2 = y + y, z2 = z + z;
  m[0] = 1.f - (y * y2 + z * z2); m[1] = x * y2 - w * z2;         m[2] = x * z2 + w * y2;
  m[3] = x * y2 + w * z2;         m[4] = 1.f - (x * x2 + z * z2); m[5] = y * z2 - w * x2;
  m[6] = x * z2 - w * y2;         m[7] = y * z2 + w * x2;         m[8] = 1.f - (x * x2 + y * y2);
}

bool file_exists(const std::string& p) { return std::ifstream(p).good(); }
long long file_mtime(const std::string& p) {
  std::error_code ec;
  const auto t = std::filesystem::last_write_time(p, ec);
  return ec ? 0 : (long long)t.time_since_epoch().count();
}

}  // namespace

extern "C" {

DP_EXPORT DragPoser* init_drag_poser(void) {
  try {
    return new DragPoser();
  } catch (...) {
    log_line("init_drag_poser: allocation failed");
    return nullptr;
  }
}

DP_EXPORT void set_reference_skeleton(DragPoser* d, char* bvhPath) {
  if (!d || !bvhPath) return;
  std::vector<int32_t> par;
  std::vector<float> off;
  std::string err;
  if (!parse_bvh_hierarchy(bvhPath, par, off, err)) { d->fail("set_reference_skeleton: " + err); return; }
  if ((int)par.size() != J) { d->fail("set_reference_skeleton: the pose model needs a 22-joint skeleton, got " + std::to_string(par.size())); return; }
  d->parents = par;
  d->offsets = off;
  d->num_joints = (int)par.size();
  d->have_skeleton = true;
  d->ok();
}

DP_EXPORT void load_models(DragPoser* d, char* modelPath) {
  if (!d || !modelPath) return;
  if (!d->have_skeleton) { d->fail("load_models: call set_reference_skeleton first"); return; }
  std::string dir(modelPath);
  while (dir.size() > 1 && (dir.back() == '/' || dir.back() == '\\')) dir.pop_back();
  std::string dpm = dir.size() > 4 && dir.substr(dir.size() - 4) == ".dpm" ? dir : dir + "/model.dpm";
  // The reference checkpoint layout (generator.pt, data.pt, temporal.pt: PyTorch pickles) is converted OFFLINE by
  // `python -m dragposer_b200.export_model <dir>`; this library never spawns an interpreter or a shell.  A model
  // directory that is read-only can keep its model.dpm in $DRAGPOSER_MODEL_CACHE.
  if (!file_exists(dpm)) {
    if (const char* cache = getenv("DRAGPOSER_MODEL_CACHE")) dpm = std::string(cache) + "/model.dpm";
    if (!file_exists(dpm)) {
      d->fail("load_models: " + dpm + " not found -- convert the checkpoint once with: python -m dragposer_b200.export_model \"" + dir + "\"");
      return;
    }
  }
  const std::string tpt = dir + "/temporal.pt";
  if (file_exists(tpt) && file_mtime(tpt) > file_mtime(dpm)) {
    d->fail("load_models: " + dpm + " is older than " + tpt + " -- re-run python -m dragposer_b200.export_model");
    return;
  }
  std::string err;
  bool trained = false;
  if (!read_dpm(dpm, d->model, trained, err)) { d->fail("load_models: " + err); return; }
  if (!trained) log_line("load_models: " + dpm + " carries the random-init temporal predictor (exported with --random-temporal), not a trained temporal.pt");
  if (d->engine) { dp_engine_destroy(d->engine); d->engine = nullptr; }
  int dev = 0;
  if (const char* s = getenv("DRAGPOSER_DEVICE")) dev = atoi(s);
  if (dp_engine_create(&d->engine, dev, 1) != DP_OK) { d->fail(std::string("load_models: ") + dp_engine_last_error()); return; }
  const Model& m = d->model;
  dp_pose_model pm{m.A0.data(), m.b0.data(), m.A1.data(), m.b1.data(), m.A2.data(), m.b2.data(), m.mean_q.data(), m.std_q.data(),
                   m.mean_d.data(), m.std_d.data(), d->parents.data(), d->offsets.data()};
  if (dp_engine_set_pose_model(d->engine, &pm) != DP_OK ||
      dp_engine_set_temporal_model(d->engine, m.temporal.data(), m.temporal.size(), m.means_latent.data(), m.stds_latent.data()) != DP_OK) {
    d->fail(std::string("load_models: ") + dp_engine_last_error());
    return;
  }
  d->have_model = true;
  d->ok();
}

DP_EXPORT void set_mask_and_weights(DragPoser* d, float* mask, float2* weights) {
  if (!d || !mask || !weights) return;
  if (!d->have_skeleton) { d->fail("set_mask_and_weights: call set_reference_skeleton first"); return; }
  std::vector<int32_t> joints;
  std::vector<float> w;
  for (int j = 0; j < d->num_joints; ++j)
    if (mask[j] != 0.0f) {  // torch.nonzero(mask) (run_drag.py:73)
      joints.push_back(j);
      w.push_back(weights[j].x);
      w.push_back(weights[j].y);
    }
  if (joints.size() < 2) { d->fail("set_mask_and_weights: at least two trackers are required (the reference cannot run one)"); return; }
  d->ee_joints = joints;
  d->ee_weights = w;
  d->num_ee = (int)joints.size();
  d->ok();
}

DP_EXPORT void init_drag_model(DragPoser* d, float3 pos, quaternion rot) {
  if (!d) return;
  if (!d->have_model) { d->fail("init_drag_model: call load_models first"); return; }
  float latent[DP_LATENT];
  if (d->have_latent_override) {
    memcpy(latent, d->latent0, sizeof(latent));
  } else {
    float mu[24], logvar[24];
    encode_zero_pose(d->model, mu, logvar);
    std::normal_distribution<float> gauss(0.f, 1.f);
    for (int i = 0; i < DP_LATENT; ++i) latent[i] = mu[i] + gauss(d->rng) * std::exp(0.5f * logvar[i]);  // autoencoder.py:19-22
    memcpy(d->latent0, latent, sizeof(latent));
  }
  const float gp[3] = {pos.x, pos.y, pos.z}, gr[4] = {rot.w, rot.x, rot.y, rot.z}, heights[DP_HEIGHTS] = {0, 0, 0, 0, 0, 0};
  if (dp_engine_init_clips(d->engine, 1, latent, gp, gr, heights) != DP_OK) { d->fail(std::string("init_drag_model: ") + dp_engine_last_error()); return; }
  d->have_session = true;
  d->ok();
}

DP_EXPORT void set_optim_params(DragPoser* d, float stopEpsPos, float stopEpsRot, int maxIter, float lr) {
  if (!d) return;
  d->params.stop_eps_pos = (double)stopEpsPos;
  d->params.stop_eps_rot = (double)stopEpsRot;
  d->params.max_iter = maxIter;
  d->params.learning_rate = lr;
  d->ok();
}

DP_EXPORT void set_lambdas(DragPoser* d, float lambdaRot, float lambdaTemporal, int temporalFutureWindow) {
  if (!d) return;
  d->params.lambda_rot = lambdaRot;
  d->params.lambda_temporal = lambdaTemporal;
  d->params.temporal_future_window = temporalFutureWindow;
  d->ok();
}

DP_EXPORT void set_global_pos(DragPoser* d, float3 p) {
  if (!d) return;
  if (!d->have_session) { d->fail("set_global_pos: call init_drag_model first"); return; }
  const float gp[3] = {p.x, p.y, p.z};
  if (dp_engine_set_global_pos(d->engine, 0, 1, gp) != DP_OK) { d->fail(std::string("set_global_pos: ") + dp_engine_last_error()); return; }
  d->ok();
}

DP_EXPORT void drag_pose(DragPoser* d, int nEndEffectors, float3* targetEEPos, quaternion* targetEERot, quaternion* resultPose,
                         float3* resultGlobalPos) {
  if (!d || !targetEEPos || !targetEERot || !resultPose || !resultGlobalPos) return;
  if (!d->have_session) { d->fail("drag_pose: call init_drag_model first"); return; }
  if (nEndEffectors != d->num_ee) {  // the reference asserts (exportFunc.cpp:76)
    d->fail("drag_pose: nEndEffectors does not match the mask set by set_mask_and_weights");
    return;
  }
  const int E = nEndEffectors;
  d->tp.resize((size_t)E * 3);
  d->tr.resize((size_t)E * 9);
  d->pose.resize(DP_POSE);
  for (int e = 0; e < E; ++e) {
    d->tp[3 * e] = targetEEPos[e].x;
    d->tp[3 * e + 1] = targetEEPos[e].y;
    d->tp[3 * e + 2] = targetEEPos[e].z;
    quat_to_matrix(targetEERot[e], &d->tr[9 * e]);  // run_drag.py:136
  }
  float gp[3];
  if (dp_engine_run_frame_host(d->engine, &d->params, nullptr, d->ee_joints.data(), d->ee_weights.data(), 1, d->tp.data(), d->tr.data(), E,
                               d->pose.data(), gp) != DP_OK) {
    d->fail(std::string("drag_pose: ") + dp_engine_last_error());
    return;
  }
  // de-standardise, root-space -> parent-local quaternions (run_drag.py:161-166, train.py:409-434)
  float q[J][4];
  for (int j = 0; j < J; ++j)
    for (int i = 0; i < 4; ++i) q[j][i] = d->pose[4 * j + i] * d->model.std_q[4 * j + i] + d->model.mean_q[4 * j + i];
  for (int j = J - 1; j >= 1; --j) {
    const int p = d->parents[j];
    if (p == 0) continue;
    const float inv[4] = {q[p][0], -q[p][1], -q[p][2], -q[p][3]};
    float r[4];
    quat_mul(inv, q[j], r);
    memcpy(q[j], r, sizeof(r));
  }
  for (int j = 0; j < J; ++j) resultPose[j] = quaternion{q[j][0], q[j][1], q[j][2], q[j][3]};
  resultGlobalPos[0] = float3{gp[0], gp[1], gp[2]};
  d->ok();
}

DP_EXPORT void destroy_drag_poser(DragPoser* d) {
  if (!d) return;
  if (d->engine) dp_engine_destroy(d->engine);
  delete d;
}

DP_EXPORT int dp_last_status(const DragPoser* d) { return d ? d->status : -1; }
DP_EXPORT const char* dp_last_message(const DragPoser* d) { return d ? d->message.c_str() : "null session"; }
DP_EXPORT int dp_get_num_joints(const DragPoser* d) { return d ? d->num_joints : 0; }
DP_EXPORT int dp_get_last_iterations(DragPoser* d) {
  int32_t it = -1;
  if (!d || !d->have_session || dp_engine_get_frame_stats(d->engine, &it, nullptr) != DP_OK) return -1;
  return it;
}
DP_EXPORT void dp_set_min_loss_increment(DragPoser* d, double v) {
  if (d) d->params.min_loss_incr = v;
}
DP_EXPORT int dp_get_num_endeffectors(const DragPoser* d) { return d ? d->num_ee : 0; }
DP_EXPORT void dp_set_initial_latent(DragPoser* d, const float* latent24) {
  if (!d || !latent24) return;
  memcpy(d->latent0, latent24, sizeof(d->latent0));
  d->have_latent_override = true;
}
DP_EXPORT void dp_get_initial_latent(const DragPoser* d, float* latent24) {
  if (d && latent24) memcpy(latent24, d->latent0, sizeof(d->latent0));
}

}  // extern "C"
