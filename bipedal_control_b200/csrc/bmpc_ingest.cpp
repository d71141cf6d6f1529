// Readers for the reference's own configuration files and the model reduction that BipedalRobotInterface performs:
//   - Boost property-tree INFO subset (task.info / reference.info / gait.info): nested blocks, "[i]" lists, "(r,c)" matrix
//     entries with the optional `scaling` key, ';' and '//' comments  ([UPSTREAM] loadData::loadEigenMatrix / loadStdVector,
//     used at ocs2_bipedal_robot/src/BipedalRobotInterface.cpp:96-108, 275-276, gait/ModeSequenceTemplate.cpp:50-111)
//   - URDF subset: links (inertial), joints (origin, axis, parent, child, limit)
//   - centroidal_model::createPinocchioInterface(urdf, jointNames) [UPSTREAM] (BipedalRobotInterface.cpp:117): joints that are not
//     listed become fixed and their bodies are lumped into the parent at joint angle zero
//   - BipedalRobotInterface::initializeInputCostWeight (BipedalRobotInterface.cpp:239-269): R_joint = J^T R_v J at initialState
#include <cmath>
#include <cstring>
#include <fstream>
#include <functional>
#include <memory>
#include <sstream>
#include <stdexcept>

#include "bmpc_model.h"

namespace bmpc {

void finalize_model(HostModel& m);

namespace {

// ---------------------------------------------------------------- INFO
struct Info {
  std::vector<std::pair<std::string, std::string>> values;
  std::vector<std::pair<std::string, std::unique_ptr<Info>>> children;
  const Info* child(const std::string& k) const { for (auto& c : children) if (c.first == k) return c.second.get(); return nullptr; }
  const std::string* value(const std::string& k) const { for (auto& v : values) if (v.first == k) return &v.second; return nullptr; }
};

std::string strip_comment(std::string line) {
  size_t k = line.find(';'); if (k != std::string::npos) line = line.substr(0, k);
  k = line.find("//"); if (k != std::string::npos) line = line.substr(0, k);
  const size_t a = line.find_first_not_of(" \t\r\n");
  if (a == std::string::npos) return "";
  const size_t b = line.find_last_not_of(" \t\r\n");
  return line.substr(a, b - a + 1);
}

std::unique_ptr<Info> parse_info(const std::string& path) {
  std::ifstream fh(path);
  if (!fh) throw std::invalid_argument("[bmpc] file not found: " + path);
  auto root = std::make_unique<Info>();
  std::vector<Info*> stack{root.get()};
  std::string pending, raw;
  while (std::getline(fh, raw)) {
    std::string line = strip_comment(raw);
    while (!line.empty()) {
      if (line[0] == '{') {
        if (pending.empty()) throw std::runtime_error("[bmpc] " + path + ": '{' without key");
        auto c = std::make_unique<Info>(); Info* cp = c.get();
        // a key that was first recorded with an empty value becomes a block
        auto& vals = stack.back()->values;
        for (size_t i = 0; i < vals.size(); ++i) if (vals[i].first == pending && vals[i].second.empty()) { vals.erase(vals.begin() + i); break; }
        stack.back()->children.emplace_back(pending, std::move(c)); stack.push_back(cp); pending.clear();
        line = strip_comment(line.substr(1)); continue;
      }
      if (line[0] == '}') { if (stack.size() > 1) stack.pop_back(); pending.clear(); line = strip_comment(line.substr(1)); continue; }
      std::istringstream is(line);
      std::string key, val; is >> key; is >> val;
      if (!val.empty() && val[0] == '{') { pending = key; line = line.substr(line.find('{')); continue; }
      stack.back()->values.emplace_back(key, val); pending = key; line.clear();
    }
  }
  return root;
}

const Info* walk(const Info* n, const std::string& dotted) {
  std::istringstream is(dotted); std::string k;
  while (n && std::getline(is, k, '.')) n = n->child(k);
  return n;
}
std::string get_str(const Info* root, const std::string& dotted, bool required = true) {
  const size_t p = dotted.rfind('.');
  const Info* n = p == std::string::npos ? root : walk(root, dotted.substr(0, p));
  const std::string key = p == std::string::npos ? dotted : dotted.substr(p + 1);
  const std::string* v = n ? n->value(key) : nullptr;
  if (!v) { if (required) throw std::runtime_error("[bmpc] missing INFO key " + dotted); return ""; }
  return *v;
}
double get_num(const Info* root, const std::string& dotted) { return std::strtod(get_str(root, dotted).c_str(), nullptr); }
std::vector<std::string> get_list(const Info* root, const std::string& dotted) {
  std::vector<std::string> out; const Info* n = walk(root, dotted);
  if (!n) return out;
  for (int i = 0;; ++i) { const std::string* v = n->value("[" + std::to_string(i) + "]"); if (!v) break; out.push_back(*v); }
  return out;
}
std::vector<double> get_matrix(const Info* root, const std::string& dotted, int rows, int cols) {
  std::vector<double> m((size_t)rows * cols, 0.0); const Info* n = walk(root, dotted);
  if (!n) return m;
  double scaling = 1.0; if (const std::string* s = n->value("scaling")) scaling = std::strtod(s->c_str(), nullptr);
  for (auto& kv : n->values) { int r, c; if (std::sscanf(kv.first.c_str(), "(%d,%d)", &r, &c) == 2 && r < rows && c < cols) m[(size_t)r * cols + c] = std::strtod(kv.second.c_str(), nullptr) * scaling; }
  return m;
}
int mode_id(const std::string& s) { if (s == "FLY") return 0; if (s == "LF") return 1; if (s == "RF") return 2; if (s == "STANCE") return 3; throw std::runtime_error("[bmpc] unknown mode " + s); }

// ---------------------------------------------------------------- small linear algebra
struct V3 { double v[3]; };
struct M3 { double m[9]; };
M3 eye3() { return M3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
M3 mul(const M3& a, const M3& b) { M3 c{}; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += a.m[3 * i + k] * b.m[3 * k + j]; c.m[3 * i + j] = s; } return c; }
M3 tr(const M3& a) { M3 c{}; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) c.m[3 * i + j] = a.m[3 * j + i]; return c; }
V3 mul(const M3& a, const V3& x) { V3 y{}; for (int i = 0; i < 3; ++i) y.v[i] = a.m[3 * i] * x.v[0] + a.m[3 * i + 1] * x.v[1] + a.m[3 * i + 2] * x.v[2]; return y; }
V3 add(const V3& a, const V3& b) { return V3{{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]}}; }
V3 sub(const V3& a, const V3& b) { return V3{{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]}}; }
V3 cross(const V3& a, const V3& b) { return V3{{a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2], a.v[0] * b.v[1] - a.v[1] * b.v[0]}}; }
M3 rpy(const V3& e) {
  const double cr = std::cos(e.v[0]), sr = std::sin(e.v[0]), cp = std::cos(e.v[1]), sp = std::sin(e.v[1]), cy = std::cos(e.v[2]), sy = std::sin(e.v[2]);
  const M3 Rx{{1, 0, 0, 0, cr, -sr, 0, sr, cr}}, Ry{{cp, 0, sp, 0, 1, 0, -sp, 0, cp}}, Rz{{cy, -sy, 0, sy, cy, 0, 0, 0, 1}};
  return mul(Rz, mul(Ry, Rx));
}
M3 rot_axis(const V3& a, double ang) {
  const double s = std::sin(ang), c = std::cos(ang), t = 1 - c;
  return M3{{c + t * a.v[0] * a.v[0], t * a.v[0] * a.v[1] - s * a.v[2], t * a.v[0] * a.v[2] + s * a.v[1],
             t * a.v[0] * a.v[1] + s * a.v[2], c + t * a.v[1] * a.v[1], t * a.v[1] * a.v[2] - s * a.v[0],
             t * a.v[0] * a.v[2] - s * a.v[1], t * a.v[1] * a.v[2] + s * a.v[0], c + t * a.v[2] * a.v[2]}};
}
// inertia about the frame origin from (m, c, Ic): Ic + m (|c|^2 I - c c^T)
M3 shift_inertia(const M3& Ic, double m, const V3& c, double sign) {
  M3 r = Ic; const double cc = c.v[0] * c.v[0] + c.v[1] * c.v[1] + c.v[2] * c.v[2];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[3 * i + j] += sign * m * ((i == j ? cc : 0.0) - c.v[i] * c.v[j]);
  return r;
}

// ---------------------------------------------------------------- URDF
struct XmlNode { std::string tag; std::map<std::string, std::string> attr; std::vector<std::unique_ptr<XmlNode>> kids; const XmlNode* kid(const std::string& t) const { for (auto& k : kids) if (k->tag == t) return k.get(); return nullptr; } };

std::unique_ptr<XmlNode> parse_xml(const std::string& path) {
  std::ifstream fh(path);
  if (!fh) throw std::invalid_argument("[bmpc] URDF file not found: " + path);
  std::stringstream ss; ss << fh.rdbuf(); const std::string s = ss.str();
  auto root = std::make_unique<XmlNode>(); root->tag = "#root";
  std::vector<XmlNode*> stack{root.get()};
  size_t i = 0;
  while ((i = s.find('<', i)) != std::string::npos) {
    if (s.compare(i, 4, "<!--") == 0) { i = s.find("-->", i); if (i == std::string::npos) break; i += 3; continue; }
    if (s[i + 1] == '?' || s[i + 1] == '!') { i = s.find('>', i); if (i == std::string::npos) break; ++i; continue; }
    const size_t e = s.find('>', i);
    if (e == std::string::npos) break;
    std::string body = s.substr(i + 1, e - i - 1);
    i = e + 1;
    if (!body.empty() && body[0] == '/') { if (stack.size() > 1) stack.pop_back(); continue; }
    const bool selfclose = !body.empty() && body.back() == '/';
    if (selfclose) body.pop_back();
    auto n = std::make_unique<XmlNode>();
    size_t p = 0;
    while (p < body.size() && !std::isspace((unsigned char)body[p])) ++p;
    n->tag = body.substr(0, p);
    while (p < body.size()) {
      while (p < body.size() && std::isspace((unsigned char)body[p])) ++p;
      const size_t eq = body.find('=', p);
      if (eq == std::string::npos) break;
      std::string key = body.substr(p, eq - p);
      while (!key.empty() && std::isspace((unsigned char)key.back())) key.pop_back();
      const size_t q0 = body.find_first_of("\"'", eq);
      if (q0 == std::string::npos) break;
      const size_t q1 = body.find(body[q0], q0 + 1);
      if (q1 == std::string::npos) break;
      n->attr[key] = body.substr(q0 + 1, q1 - q0 - 1);
      p = q1 + 1;
    }
    XmlNode* np = n.get();
    stack.back()->kids.push_back(std::move(n));
    if (!selfclose) stack.push_back(np);
  }
  return root;
}
V3 parse_v3(const std::string& s, const V3& def) { if (s.empty()) return def; V3 v{}; std::istringstream is(s); is >> v.v[0] >> v.v[1] >> v.v[2]; return v; }
std::string attr(const XmlNode* n, const std::string& k) { if (!n) return ""; auto it = n->attr.find(k); return it == n->attr.end() ? "" : it->second; }

struct ULink { double mass = 0; V3 com{}; M3 I{}; };
struct UJoint { std::string name, type, parent, child; V3 xyz{}; M3 R = eye3(); V3 axis{{1, 0, 0}}; double lo = -1e9, hi = 1e9; };
struct Lump { double m = 0; V3 mc{}; M3 Io{}; void add(double mm, const V3& c, const M3& Ic) { m += mm; for (int i = 0; i < 3; ++i) mc.v[i] += mm * c.v[i]; const M3 s = shift_inertia(Ic, mm, c, +1.0); for (int i = 0; i < 9; ++i) Io.m[i] += s.m[i]; } };

}  // namespace

HostModel load_reference_files(const std::string& task_file, const std::string& reference_file, const std::string& gait_file, const std::string& urdf_file) {
  auto task = parse_info(task_file);
  auto ref = parse_info(reference_file);
  std::unique_ptr<Info> gait = gait_file.empty() ? std::make_unique<Info>() : parse_info(gait_file);
  const std::vector<std::string> joint_names = get_list(task.get(), "model_settings.jointNames");
  const std::vector<std::string> contact_names = get_list(task.get(), "model_settings.contactNames3DoF");
  if (joint_names.empty() || contact_names.size() != NCON) throw std::invalid_argument("[bmpc] task file must list jointNames and four contactNames3DoF");
  // ---- URDF
  auto xml = parse_xml(urdf_file);
  const XmlNode* robot = xml->kid("robot");
  if (!robot) throw std::invalid_argument("[bmpc] URDF has no <robot> element");
  std::map<std::string, ULink> links; std::vector<std::string> link_order;
  std::vector<UJoint> joints;
  for (auto& k : robot->kids) {
    if (k->tag == "link") {
      ULink L; const XmlNode* in = k->kid("inertial");
      if (in) {
        const XmlNode* o = in->kid("origin");
        L.com = parse_v3(attr(o, "xyz"), V3{}); const M3 Ri = rpy(parse_v3(attr(o, "rpy"), V3{}));
        L.mass = std::strtod(attr(in->kid("mass"), "value").c_str(), nullptr);
        const XmlNode* it = in->kid("inertia");
        auto g = [&](const char* a) { return std::strtod(attr(it, a).c_str(), nullptr); };
        const M3 I0{{g("ixx"), g("ixy"), g("ixz"), g("ixy"), g("iyy"), g("iyz"), g("ixz"), g("iyz"), g("izz")}};
        L.I = mul(Ri, mul(I0, tr(Ri)));
      }
      links[attr(k.get(), "name")] = L; link_order.push_back(attr(k.get(), "name"));
    } else if (k->tag == "joint") {
      UJoint J; J.name = attr(k.get(), "name"); J.type = attr(k.get(), "type");
      if (J.type == "floating") continue;
      const XmlNode* o = k->kid("origin");
      J.xyz = parse_v3(attr(o, "xyz"), V3{}); J.R = rpy(parse_v3(attr(o, "rpy"), V3{}));
      J.parent = attr(k->kid("parent"), "link"); J.child = attr(k->kid("child"), "link");
      J.axis = parse_v3(attr(k->kid("axis"), "xyz"), V3{{1, 0, 0}});
      const XmlNode* lim = k->kid("limit");
      if (lim) { if (!attr(lim, "lower").empty()) J.lo = std::strtod(attr(lim, "lower").c_str(), nullptr); if (!attr(lim, "upper").empty()) J.hi = std::strtod(attr(lim, "upper").c_str(), nullptr); }
      joints.push_back(J);
    }
  }
  std::map<std::string, std::vector<int>> children; std::map<std::string, bool> is_child;
  for (size_t i = 0; i < joints.size(); ++i) { children[joints[i].parent].push_back((int)i); is_child[joints[i].child] = true; }
  std::string root_link;
  for (auto& n : link_order) if (!is_child[n] && (children.count(n) || links[n].mass > 0)) { root_link = n; break; }
  if (root_link.empty()) throw std::invalid_argument("[bmpc] URDF root link not found");
  // ---- reduction: movable joints in DFS order
  struct Mov { std::string name; int parent; M3 R; V3 p; V3 axis; Lump body; double lo, hi; };
  std::vector<Mov> mov; Lump base; std::map<std::string, std::pair<int, V3>> contact_at;
  // bmpc extension: `contact_frames { name { parent <link>  x ..  y ..  z .. } }` defines contact points that are not URDF links
  struct CFrame { std::string name, parent; V3 xyz; };
  std::vector<CFrame> cframes;
  if (const Info* cf = task->child("contact_frames"))
    for (auto& c : cf->children) {
      const std::string* par = c.second->value("parent"); const std::string *sx = c.second->value("x"), *sy = c.second->value("y"), *sz = c.second->value("z");
      if (!par || !sx || !sy || !sz) throw std::invalid_argument("[bmpc] contact_frames." + c.first + " needs parent, x, y, z");
      cframes.push_back({c.first, *par, V3{{std::strtod(sx->c_str(), nullptr), std::strtod(sy->c_str(), nullptr), std::strtod(sz->c_str(), nullptr)}}});
    }
  auto is_listed = [&](const std::string& n) { for (auto& s : joint_names) if (s == n) return true; return false; };
  std::function<void(const std::string&, int, const M3&, const V3&)> visit = [&](const std::string& link, int mi, const M3& Racc, const V3& pacc) {
    const ULink& L = links[link];
    if (L.mass > 0) { Lump& body = mi < 0 ? base : mov[mi].body; body.add(L.mass, add(mul(Racc, L.com), pacc), mul(Racc, mul(L.I, tr(Racc)))); }
    for (auto& c : contact_names) if (c == link) contact_at[link] = {mi, pacc};
    for (auto& cfr : cframes) if (cfr.parent == link) contact_at[cfr.name] = {mi, add(mul(Racc, cfr.xyz), pacc)};
    auto it = children.find(link);
    if (it == children.end()) return;
    for (int ji : it->second) {
      const UJoint& J = joints[ji];
      const M3 Rj = mul(Racc, J.R); const V3 pj = add(mul(Racc, J.xyz), pacc);
      if (is_listed(J.name)) {
        if (J.type != "revolute" && J.type != "continuous") throw std::invalid_argument("[bmpc] listed joint " + J.name + " is not revolute");
        double nrm = std::sqrt(J.axis.v[0] * J.axis.v[0] + J.axis.v[1] * J.axis.v[1] + J.axis.v[2] * J.axis.v[2]);
        Mov m{J.name, mi, Rj, pj, V3{{J.axis.v[0] / nrm, J.axis.v[1] / nrm, J.axis.v[2] / nrm}}, Lump{}, J.lo, J.hi};
        mov.push_back(m);
        visit(J.child, (int)mov.size() - 1, eye3(), V3{});
      } else visit(J.child, mi, Rj, pj);
    }
  };
  visit(root_link, -1, eye3(), V3{});
  if (mov.size() != joint_names.size()) throw std::invalid_argument("[bmpc] not every listed joint was found in the URDF");
  for (size_t j = 0; j < mov.size(); ++j) if (mov[j].name != joint_names[j]) throw std::invalid_argument("[bmpc] joint order in the task file differs from the URDF tree order");
  HostModel m; DevModel& d = m.dev; std::memset(&d, 0, sizeof(d));
  m.name = attr(robot, "name"); m.nj = (int)mov.size();
  if (m.nj > MAXJ) throw std::invalid_argument("[bmpc] too many leg joints");
  auto finish = [](const Lump& b, double& mass, double* com, double* I) {
    mass = b.m; V3 c{}; for (int i = 0; i < 3; ++i) { c.v[i] = b.m > 0 ? b.mc.v[i] / b.m : 0.0; com[i] = c.v[i]; }
    const M3 Ic = shift_inertia(b.Io, b.m, c, -1.0); for (int i = 0; i < 9; ++i) I[i] = Ic.m[i];
  };
  finish(base, d.base_mass, d.base_com, d.base_inertia);
  d.total_mass = d.base_mass;
  for (int j = 0; j < m.nj; ++j) {
    m.joint_names.push_back(mov[j].name); m.joint_parent.push_back(mov[j].parent); m.joint_lo.push_back(mov[j].lo); m.joint_hi.push_back(mov[j].hi);
    for (int i = 0; i < 9; ++i) d.Rj[j][i] = mov[j].R.m[i];
    for (int i = 0; i < 3; ++i) { d.pj[j][i] = mov[j].p.v[i]; d.axis[j][i] = mov[j].axis.v[i]; }
    finish(mov[j].body, d.mass[j], d.com[j], d.inertia[j]);
    d.total_mass += d.mass[j];
  }
  for (int c = 0; c < NCON; ++c) {
    auto it = contact_at.find(contact_names[c]);
    if (it == contact_at.end()) throw std::invalid_argument("[bmpc] contact link " + contact_names[c] + " not found in the URDF");
    m.contact_names.push_back(contact_names[c]); m.contact_parent.push_back(it->second.first);
    for (int i = 0; i < 3; ++i) d.coff[c][i] = it->second.second.v[i];
  }
  // ---- problem data
  const int nx = 12 + m.nj;
  m.initial_state = get_matrix(task.get(), "initialState", nx, 1);
  const std::vector<double> Q = get_matrix(task.get(), "Q", nx, nx);
  for (int i = 0; i < nx; ++i) for (int j = 0; j < nx; ++j) if (i != j && Q[(size_t)i * nx + j] != 0.0) throw std::invalid_argument("[bmpc] only diagonal Q is supported");
  for (int i = 0; i < nx; ++i) d.Qdiag[i] = Q[(size_t)i * nx + i];
  const int nt = 6 * NCON;
  const std::vector<double> Rt = get_matrix(task.get(), "R", nt, nt);
  m.R_taskspace_diag.resize(nt);
  for (int i = 0; i < nt; ++i) { for (int j = 0; j < nt; ++j) if (i != j && Rt[(size_t)i * nt + j] != 0.0) throw std::invalid_argument("[bmpc] only diagonal task-space R is supported"); m.R_taskspace_diag[i] = Rt[(size_t)i * nt + i]; }
  for (int i = 0; i < 12; ++i) d.Rforce[i] = m.R_taskspace_diag[i];
  {  // R_joint = J^T R_v J with the contact Jacobians w.r.t. the leg joints at initialState (BipedalRobotInterface.cpp:247-269)
    const double* q = m.initial_state.data() + 6;
    const double cz = std::cos(q[3]), sz = std::sin(q[3]), cy = std::cos(q[4]), sy = std::sin(q[4]), cx = std::cos(q[5]), sx = std::sin(q[5]);
    const M3 Rb{{cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx, sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx, -sy, cy * sx, cy * cx}};
    const V3 pb{{q[0], q[1], q[2]}};
    std::vector<M3> Rw(m.nj); std::vector<V3> pw(m.nj), aw(m.nj);
    for (int j = 0; j < m.nj; ++j) {
      const M3& Rp = mov[j].parent < 0 ? Rb : Rw[mov[j].parent]; const V3& pp = mov[j].parent < 0 ? pb : pw[mov[j].parent];
      const M3 Rfix = mul(Rp, mov[j].R);
      pw[j] = add(mul(Rp, mov[j].p), pp); aw[j] = mul(Rfix, mov[j].axis); Rw[j] = mul(Rfix, rot_axis(mov[j].axis, q[6 + j]));
    }
    std::vector<double> J((size_t)3 * NCON * m.nj, 0.0);
    for (int c = 0; c < NCON; ++c) {
      const int par = m.contact_parent[c];
      const V3 off{{d.coff[c][0], d.coff[c][1], d.coff[c][2]}};
      const V3 pc = par < 0 ? add(mul(Rb, off), pb) : add(mul(Rw[par], off), pw[par]);
      for (int k = par; k >= 0; k = mov[k].parent) { const V3 col = cross(aw[k], sub(pc, pw[k])); for (int r = 0; r < 3; ++r) J[(size_t)(3 * c + r) * m.nj + k] = col.v[r]; }
    }
    for (int a = 0; a < m.nj; ++a) for (int b = 0; b < m.nj; ++b) { double s = 0; for (int r = 0; r < 3 * NCON; ++r) s += J[(size_t)r * m.nj + a] * m.R_taskspace_diag[3 * NCON + r] * J[(size_t)r * m.nj + b]; d.Rjoint[a * m.nj + b] = s; }
  }
  m.default_joint_state = get_matrix(ref.get(), "defaultJointState", m.nj, 1);
  m.com_height = get_num(ref.get(), "comHeight"); m.target_disp_vel = get_num(ref.get(), "targetDisplacementVelocity"); m.target_rot_vel = get_num(ref.get(), "targetRotationVelocity");
  d.mu_f = get_num(task.get(), "frictionConeSoftConstraint.frictionCoefficient"); d.bar_mu = get_num(task.get(), "frictionConeSoftConstraint.mu"); d.bar_delta = get_num(task.get(), "frictionConeSoftConstraint.delta");
  d.fr_reg = 25.0; d.fr_grip = 0.0; d.fr_shift = 1e-6;   // FrictionConeConstraint.h:66-67 constructor defaults
  d.gain = get_num(task.get(), "model_settings.positionErrorGain"); m.phase_transition_stance_time = get_num(task.get(), "model_settings.phaseTransitionStanceTime");
  d.liftoff_vel = get_num(task.get(), "swing_trajectory_config.liftOffVelocity"); d.touchdown_vel = get_num(task.get(), "swing_trajectory_config.touchDownVelocity");
  d.swing_height = get_num(task.get(), "swing_trajectory_config.swingHeight"); d.swing_time_scale = get_num(task.get(), "swing_trajectory_config.swingTimeScale");
  m.sqp_dt = get_num(task.get(), "sqp.dt"); m.sqp_iterations = (int)get_num(task.get(), "sqp.sqpIteration"); d.delta_tol = get_num(task.get(), "sqp.deltaTol");
  d.g_max = get_num(task.get(), "sqp.g_max"); d.g_min = get_num(task.get(), "sqp.g_min");
  m.time_horizon = get_num(task.get(), "mpc.timeHorizon"); m.mpc_frequency = get_num(task.get(), "mpc.mpcDesiredFrequency");
  for (auto& s : get_list(ref.get(), "initialModeSchedule.modeSequence")) m.init_modes.push_back(mode_id(s));
  for (auto& s : get_list(ref.get(), "initialModeSchedule.eventTimes")) m.init_events.push_back(std::strtod(s.c_str(), nullptr));
  for (auto& s : get_list(ref.get(), "defaultModeSequenceTemplate.modeSequence")) m.default_template.modes.push_back(mode_id(s));
  for (auto& s : get_list(ref.get(), "defaultModeSequenceTemplate.switchingTimes")) m.default_template.times.push_back(std::strtod(s.c_str(), nullptr));
  m.default_template.name = "default";
  for (auto& g : get_list(gait.get(), "list")) {
    GaitTemplate t; t.name = g;
    for (auto& s : get_list(gait.get(), g + ".modeSequence")) t.modes.push_back(mode_id(s));
    for (auto& s : get_list(gait.get(), g + ".switchingTimes")) t.times.push_back(std::strtod(s.c_str(), nullptr));
    m.gaits.push_back(t);
  }
  finalize_model(m);
  return m;
}

}  // namespace bmpc
