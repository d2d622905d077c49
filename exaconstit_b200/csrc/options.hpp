// Input side of the application driver: the subset of TOML the reference's option files use and
// an ExaOptions mirror with the reference's key names, defaults and abort messages
//   option tables and defaults      src/option_parser.cpp:26-724, src/option_parser.hpp:60-230
//   text inputs                     src/mechanics_driver.cpp:193-216 (custom dt), 430-515 (props / state / ori),
//                                   259-281 (grain map)
// The reference parses with the third-party toml11 library; this is an independent reader for the constructs those
// files contain: comments, [tables] and [sub.tables], strings, integers, floats, booleans and (nested, multi-line)
// arrays.  Inline tables, dotted keys, dates and multi-line strings are not used by the reference's files and are
// rejected.
#pragma once
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace exaopt {

struct Abort : std::runtime_error { using std::runtime_error::runtime_error; };  // stands in for MFEM_ABORT

struct TVal {
  enum Type { STR, INT, FLT, BOOL, ARR } t = INT;
  std::string s;
  double d = 0.0;
  long i = 0;
  bool b = false;
  std::vector<TVal> a;
};

class Toml {
 public:
  std::map<std::string, TVal> vals;  // "Table.Sub.key"
  std::set<std::string> tables;

  static Toml parse_file(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw Abort("Cannot open options file: " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return parse(ss.str());
  }

  static Toml parse(const std::string& text) {
    Toml doc;
    // strip comments (outside strings), keep newlines
    std::string src;
    {
      bool in_str = false;
      char q = 0;
      for (size_t k = 0; k < text.size(); ++k) {
        const char c = text[k];
        if (in_str) {
          src += c;
          if (c == q) in_str = false;
        } else if (c == '"' || c == '\'') {
          in_str = true; q = c; src += c;
        } else if (c == '#') {
          while (k < text.size() && text[k] != '\n') ++k;
          src += '\n';
        } else src += c;
      }
    }
    std::string table;
    size_t p = 0;
    int line = 1;
    auto skip_ws = [&](bool newlines) {
      while (p < src.size() && (src[p] == ' ' || src[p] == '\t' || src[p] == '\r' || (newlines && src[p] == '\n'))) {
        if (src[p] == '\n') ++line;
        ++p;
      }
    };
    while (true) {
      skip_ws(true);
      if (p >= src.size()) break;
      if (src[p] == '[') {
        const size_t e = src.find(']', p);
        if (e == std::string::npos || src.compare(p, 2, "[[") == 0) throw Abort("options: bad table header at line " + std::to_string(line));
        table = trim(src.substr(p + 1, e - p - 1));
        std::string acc;
        std::stringstream parts(table);
        std::string part;
        while (std::getline(parts, part, '.')) { acc = acc.empty() ? trim(part) : acc + "." + trim(part); doc.tables.insert(acc); }
        p = e + 1;
        continue;
      }
      const size_t eq = src.find('=', p);
      const size_t nl = src.find('\n', p);
      if (eq == std::string::npos || (nl != std::string::npos && nl < eq)) throw Abort("options: expected key = value at line " + std::to_string(line));
      const std::string key = trim(src.substr(p, eq - p));
      if (key.empty() || key.find('.') != std::string::npos) throw Abort("options: unsupported key '" + key + "' at line " + std::to_string(line));
      p = eq + 1;
      skip_ws(false);
      TVal v = parse_value(src, p, line);
      doc.vals[table.empty() ? key : table + "." + key] = v;
      skip_ws(false);
      if (p < src.size() && src[p] != '\n') throw Abort("options: trailing characters after value of '" + key + "' at line " + std::to_string(line));
    }
    return doc;
  }

  bool contains(const std::string& key) const { return vals.count(key) != 0; }
  bool has_table(const std::string& t) const { return tables.count(t) != 0; }
  const TVal* get(const std::string& key) const {
    auto it = vals.find(key);
    return it == vals.end() ? nullptr : &it->second;
  }
  std::string str_or(const std::string& key, const std::string& dflt) const {
    const TVal* v = get(key);
    return (v && v->t == TVal::STR) ? v->s : dflt;
  }
  long int_or(const std::string& key, long dflt) const {
    const TVal* v = get(key);
    return (v && v->t == TVal::INT) ? v->i : dflt;
  }
  double num_or(const std::string& key, double dflt) const {  // integers are accepted where a float is expected
    const TVal* v = get(key);
    if (v && v->t == TVal::FLT) return v->d;
    if (v && v->t == TVal::INT) return (double)v->i;
    return dflt;
  }
  bool bool_or(const std::string& key, bool dflt) const {
    const TVal* v = get(key);
    return (v && v->t == TVal::BOOL) ? v->b : dflt;
  }
  static double as_num(const TVal& v) {
    if (v.t == TVal::FLT) return v.d;
    if (v.t == TVal::INT) return (double)v.i;
    throw Abort("options: expected a number");
  }
  // flattens arbitrarily nested numeric arrays
  static void flatten(const TVal& v, std::vector<double>& out) {
    if (v.t == TVal::ARR) { for (const TVal& e : v.a) flatten(e, out); }
    else out.push_back(as_num(v));
  }

 private:
  static std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) ++a;
    while (b > a && isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
  }
  static TVal parse_value(const std::string& src, size_t& p, int& line) {
    TVal v;
    if (p >= src.size()) throw Abort("options: missing value at line " + std::to_string(line));
    const char c = src[p];
    if (c == '"' || c == '\'') {
      const size_t e = src.find(c, p + 1);
      if (e == std::string::npos) throw Abort("options: unterminated string at line " + std::to_string(line));
      v.t = TVal::STR;
      v.s = src.substr(p + 1, e - p - 1);
      p = e + 1;
      return v;
    }
    if (c == '[') {
      v.t = TVal::ARR;
      ++p;
      while (true) {
        while (p < src.size() && (isspace((unsigned char)src[p]) || src[p] == ',')) { if (src[p] == '\n') ++line; ++p; }
        if (p >= src.size()) throw Abort("options: unterminated array");
        if (src[p] == ']') { ++p; break; }
        v.a.push_back(parse_value(src, p, line));
      }
      return v;
    }
    if (c == '{') throw Abort("options: inline tables are not supported (line " + std::to_string(line) + ")");
    size_t e = p;
    while (e < src.size() && !isspace((unsigned char)src[e]) && src[e] != ',' && src[e] != ']') ++e;
    const std::string tok = src.substr(p, e - p);
    p = e;
    if (tok == "true" || tok == "false") { v.t = TVal::BOOL; v.b = tok == "true"; return v; }
    std::string num;
    for (char ch : tok) if (ch != '_') num += ch;
    char* endp = nullptr;
    const bool is_float = num.find_first_of(".eE") != std::string::npos || num == "inf" || num == "nan";
    if (is_float) {
      v.t = TVal::FLT;
      v.d = std::strtod(num.c_str(), &endp);
    } else {
      v.t = TVal::INT;
      v.i = std::strtol(num.c_str(), &endp, 10);
    }
    if (num.empty() || !endp || *endp != '\0') throw Abort("options: cannot parse value '" + tok + "' at line " + std::to_string(line));
    return v;
  }
};

enum class XtalType { FCC = 0, BCC = 1, HCP = 2, NOTYPE };
enum class SlipType { POWERVOCE = 0, POWERVOCENL = 1, MTSDD = 2, NOTYPE };
enum class Assembly { FULL, PA, EA };
enum class IntegrationType { FULL = 0, BBAR = 1 };
enum class NLSolver { NR = 0, NRLS = 1 };
enum class KrylovSolver { GMRES, PCG, MINRES };

// One essential-BC set, active from `step` on (BCManager maps keyed by update step, src/option_parser.cpp:146-338)
struct BCSet {
  int step = 1;
  std::vector<int> ids, comps;   // comps < 0: velocity-gradient attribute
  std::vector<double> vals;      // 3 per id
  std::vector<double> vgrad;     // 9 (row-major) or empty
};

// Mirror of the reference's ExaOptions (src/option_parser.hpp:18-230): same member names where they exist there.
struct ExaOptions {
  std::string floc;
  // Properties
  double temp_k = 298.0;
  std::string props_file = "props.txt", state_file = "state.txt", ori_file = "ori.txt", grain_map = "grain_map.txt";
  int nProps = 1, numStateVars = 1, ngrains = 0, grain_custom_stride = 1, grain_statevar_offset = -1;
  std::string ori_type = "euler";
  // BCs
  bool changing_bcs = false, constant_strain_rate = false;
  std::vector<int> updateStep{1};
  std::vector<BCSet> bcs;
  bool vgrad_origin_flag = false;
  // Model
  std::string mech_type;
  bool cp = false;
  XtalType xtal_type = XtalType::NOTYPE;
  SlipType slip_type = SlipType::NOTYPE;
  // Time
  bool dt_cust = false, dt_auto = false;
  double dt = 1.0, dt_min = 1.0, dt_scale = 0.25, t_final = 1.0;
  int nsteps = 1;
  std::string dt_file = "custom_dt.txt";
  std::vector<double> cust_dt;
  // Visualizations
  int vis_steps = 1;
  std::string avg_stress_fname = "avg_stress.txt", avg_pl_work_fname = "avg_pl_work.txt",
              avg_def_grad_fname = "avg_def_grad.txt", avg_dp_tensor_fname = "avg_dp_tensor.txt";
  bool additional_avgs = false;
  // Solvers
  Assembly assembly = Assembly::FULL;
  std::string rtmodel = "CPU";
  NLSolver nl_solver = NLSolver::NR;
  IntegrationType integ_type = IntegrationType::FULL;
  double newton_rel_tol = 1.0e-5, newton_abs_tol = 1.0e-10;
  int newton_iter = 25;
  double krylov_rel_tol = 1.0e-10, krylov_abs_tol = 1.0e-30;
  int krylov_iter = 200;
  KrylovSolver solver = KrylovSolver::GMRES;
  // Mesh
  int ser_ref_levels = 0, par_ref_levels = 0, order = 1;
  std::string mesh_type = "other", mesh_file;
  double mxyz[3] = {1.0, 1.0, 1.0};
  int nxyz[3] = {1, 1, 1};

  explicit ExaOptions(const std::string& f) : floc(f) {}

  static std::string lower(std::string s) { for (char& c : s) c = (char)tolower((unsigned char)c); return s; }

  void parse_options() {
    const Toml t = Toml::parse_file(floc);
    get_properties(t);
    get_bcs(t);
    get_model(t);
    get_time_steps(t);
    get_visualizations(t);
    get_solvers(t);
    get_mesh(t);
  }

 private:
  void get_properties(const Toml& t) {
    temp_k = t.num_or("Properties.temperature", 298.0);
    if (temp_k <= 0.0) throw Abort("Properties.temperature is given in Kelvins and therefore can't be less than 0");
    if (t.has_table("Properties.Matl_Props")) {
      props_file = t.str_or("Properties.Matl_Props.floc", "props.txt");
      nProps = (int)t.int_or("Properties.Matl_Props.num_props", 1);
    } else throw Abort("Properties.Matl_Props table was not provided in toml file");
    if (t.has_table("Properties.State_Vars")) {
      numStateVars = (int)t.int_or("Properties.State_Vars.num_vars", 1);
      state_file = t.str_or("Properties.State_Vars.floc", "state.txt");
    } else throw Abort("Properties.State_Vars table was not provided in toml file");
    if (t.has_table("Properties.Grain")) {
      grain_statevar_offset = (int)t.int_or("Properties.Grain.ori_state_var_loc", -1);
      grain_custom_stride = (int)t.int_or("Properties.Grain.ori_stride", 0);
      ori_type = lower(t.str_or("Properties.Grain.ori_type", "euler"));
      if (ori_type == "quaternion") ori_type = "quat";
      if (ori_type != "quat" && ori_type != "euler" && ori_type != "custom")
        throw Abort("Properties.Grain.ori_type was not provided a valid type.");
      ngrains = (int)t.int_or("Properties.Grain.num_grains", 0);
      ori_file = t.str_or("Properties.Grain.ori_floc", "ori.txt");
      grain_map = t.str_or("Properties.Grain.grain_floc", "grain_map.txt");
    }
  }

  static std::vector<int> int_list(const TVal& v) {
    std::vector<int> out;
    for (const TVal& e : v.a) {
      if (e.t != TVal::INT) throw Abort("BCs: expected an array of integers");
      out.push_back((int)e.i);
    }
    return out;
  }

  void get_bcs(const Toml& t) {
    if (!t.has_table("BCs")) throw Abort("BCs table was not provided in toml file");
    changing_bcs = t.bool_or("BCs.changing_ess_bcs", false);
    constant_strain_rate = t.bool_or("BCs.constant_strain_rate", false);
    if (t.contains("BCs.vgrad_origin")) {
      vgrad_origin_flag = true;
      throw Abort("BCs.vgrad_origin (user-supplied velocity-gradient origin) is not supported by this driver");
    }
    updateStep = {1};
    if (const TVal* us = t.get("BCs.update_steps")) updateStep = int_list(*us);
    if (updateStep.empty()) throw Abort("BCs.update_steps was not provided any values.");
    bool has1 = false;
    for (int s : updateStep) has1 |= (s == 1);
    if (!has1) throw Abort("BCs.update_steps must contain 1 in the array");
    const TVal* ids = t.get("BCs.essential_ids");
    const TVal* comps = t.get("BCs.essential_comps");
    const TVal* vals = t.get("BCs.essential_vals");
    const TVal* vg = t.get("BCs.essential_vel_grad");
    if (!ids || ids->t != TVal::ARR || ids->a.empty()) throw Abort("BCs.essential_ids was not provided any values.");
    if (!comps || comps->t != TVal::ARR || comps->a.empty()) throw Abort("BCs.essential_comps was not provided any values.");
    const size_t nsets = changing_bcs ? updateStep.size() : 1;
    for (size_t k = 0; k < nsets; ++k) {
      BCSet b;
      b.step = changing_bcs ? updateStep[k] : 1;
      const TVal* idk = ids;
      const TVal* ck = comps;
      const TVal* vk = vals;
      const TVal* gk = vg;
      if (changing_bcs) {
        if (ids->a.size() != nsets || comps->a.size() != nsets)
          throw Abort("BCs.essential_ids / essential_comps must have one entry per BCs.update_steps entry");
        idk = &ids->a[k];
        ck = &comps->a[k];
        if (vals) { if (vals->a.size() != nsets) throw Abort("BCs.essential_vals must have one entry per update step"); vk = &vals->a[k]; }
        if (vg) { if (vg->a.size() != nsets) throw Abort("BCs.essential_vel_grad must have one entry per update step"); gk = &vg->a[k]; }
      }
      b.ids = int_list(*idk);
      b.comps = int_list(*ck);
      if (b.ids.size() != b.comps.size()) throw Abort("BCs.essential_ids and BCs.essential_comps differ in length");
      bool any_vel = false, any_vg = false;
      for (int c : b.comps) { if (c < 0) any_vg = true; else if (c > 0) any_vel = true; if (std::abs(c) > 7) throw Abort("BCs.essential_comps entries must be within -7..7"); }
      if (vk) Toml::flatten(*vk, b.vals);
      if (b.vals.empty() && any_vel) throw Abort("BCs.essential_vals was not provided any values  but a boundary requires this.");
      if (b.vals.empty()) b.vals.assign(3 * b.ids.size(), 0.0);
      if (b.vals.size() != 3 * b.ids.size()) throw Abort("BCs.essential_vals must hold 3 values per essential id");
      if (gk) Toml::flatten(*gk, b.vgrad);
      if (any_vg && b.vgrad.size() != 9) throw Abort("BCs.essential_vel_grad was not provided any values but a boundary requires this.");
      if (!any_vg) b.vgrad.clear();
      bcs.push_back(b);
    }
  }

  void get_model(const Toml& t) {
    if (!t.has_table("Model")) throw Abort("Model table was not provided in toml file");
    mech_type = lower(t.str_or("Model.mech_type", ""));
    if (mech_type == "umat") throw Abort("Model.mech_type = umat: user material libraries are outside this driver's scope (ExaCMech models only)");
    if (mech_type != "exacmech") throw Abort("Model.mech_type was not provided a valid type.");
    cp = t.bool_or("Model.cp", false);
    if (!cp) throw Abort("Model.cp needs to be set to true when using ExaCMech based models.");
    if (ori_type != "quat") throw Abort("Properties.Grain.ori_type is not set to quaternion for use with an ExaCMech model.");
    grain_statevar_offset = 9;  // ecmech::evptn::iHistLbQ
    if (!t.has_table("Model.ExaCMech")) throw Abort("The table Model.ExaCMech does not exist.");
    const std::string x = lower(t.str_or("Model.ExaCMech.xtal_type", ""));
    const std::string s = lower(t.str_or("Model.ExaCMech.slip_type", ""));
    if (x == "fcc") xtal_type = XtalType::FCC;
    else if (x == "bcc") xtal_type = XtalType::BCC;
    else if (x == "hcp") xtal_type = XtalType::HCP;
    else throw Abort("Model.ExaCMech.xtal_type was not provided a valid type.");
    int need_props = -1;
    if (s == "mts" || s == "mtsdd") {
      slip_type = SlipType::MTSDD;
      need_props = xtal_type == XtalType::HCP ? 36 : 24;
    } else if (s == "powervoce") {
      slip_type = SlipType::POWERVOCE;
      if (xtal_type == XtalType::HCP) throw Abort("Model.ExaCMech.slip_type can not be PowerVoce for HCP or BCC_112 materials.");
      need_props = 17;
    } else if (s == "powervocenl") {
      slip_type = SlipType::POWERVOCENL;
      if (xtal_type == XtalType::HCP) throw Abort("Model.ExaCMech.slip_type can not be PowerVoceNL for HCP or BCC_112 materials.");
      need_props = 18;
    } else throw Abort("Model.ExaCMech.slip_type was not provided a valid type.");
    if (nProps != need_props)
      throw Abort("Properties.Matl_Props.num_props needs " + std::to_string(need_props) + " values for the selected ExaCMech model");
    // numHist + ne + 1 - 4 (the quaternion is not part of the state file)
    const int need_state = (xtal_type == XtalType::HCP ? 40 : 28) - 4;
    if (numStateVars != need_state)
      throw Abort("Properties.State_Vars.num_vars needs " + std::to_string(need_state) + " values for this crystal type when using an "
                  "ExaCMech model. Note: the number of values for a quaternion are not included in this count.");
  }

  void get_time_steps(const Toml& t) {
    if (!t.has_table("Time")) throw Abort("Time table was not provided in toml file");
    if (t.has_table("Time.Fixed")) {
      dt_cust = false; dt_auto = false;
      dt = t.num_or("Time.Fixed.dt", 1.0);
      dt_min = dt;
      t_final = t.num_or("Time.Fixed.t_final", 1.0);
    }
    if (t.has_table("Time.Auto")) {
      if (changing_bcs) throw Abort("Automatic time stepping is currently not compatible with changing boundary conditions");
      dt_cust = false; dt_auto = true;
      dt = t.num_or("Time.Auto.dt_start", 1.0);
      dt_scale = t.num_or("Time.Auto.dt_scale", 0.25);
      if (dt_scale < 0.0 || dt_scale > 1.0) throw Abort("dt_scale for auto time stepping needs to be between 0 and 1.");
      dt_min = t.num_or("Time.Auto.dt_min", 1.0);
      t_final = t.num_or("Time.Auto.t_final", 1.0);
      dt_file = t.str_or("Time.Auto.auto_dt_file", "auto_dt_out.txt");
    }
    if (t.has_table("Time.Custom")) {
      dt_cust = true; dt_auto = false;
      nsteps = (int)t.int_or("Time.Custom.nsteps", 1);
      dt_file = t.str_or("Time.Custom.floc", "custom_dt.txt");
    }
  }

  void get_visualizations(const Toml& t) {
    vis_steps = (int)t.int_or("Visualizations.steps", 1);
    for (const char* k : {"visit", "conduit", "paraview", "adios2"})
      if (t.bool_or(std::string("Visualizations.") + k, false))
        throw Abort(std::string("Visualizations.") + k + ": field output needs MFEM data collections, which this driver does not have");
    avg_stress_fname = t.str_or("Visualizations.avg_stress_fname", "avg_stress.txt");
    additional_avgs = t.bool_or("Visualizations.additional_avgs", false);
    avg_def_grad_fname = t.str_or("Visualizations.avg_def_grad_fname", "avg_def_grad.txt");
    avg_pl_work_fname = t.str_or("Visualizations.avg_pl_work_fname", "avg_pl_work.txt");
    avg_dp_tensor_fname = t.str_or("Visualizations.avg_dp_tensor_fname", "avg_dp_tensor.txt");
  }

  void get_solvers(const Toml& t) {
    const std::string a = lower(t.str_or("Solvers.assembly", "FULL"));
    if (a == "full") assembly = Assembly::FULL;
    else if (a == "pa") assembly = Assembly::PA;
    else if (a == "ea") assembly = Assembly::EA;
    else throw Abort("Solvers.assembly was not provided a valid type.");
    rtmodel = lower(t.str_or("Solvers.rtmodel", "CPU"));
    if (rtmodel != "cpu" && rtmodel != "openmp" && rtmodel != "cuda" && rtmodel != "hip" && rtmodel != "gpu")
      throw Abort("Solvers.rtmodel was not provided a valid type.");
    if (t.has_table("Solvers.NR")) {
      const std::string s = lower(t.str_or("Solvers.NR.nl_solver", "NR"));
      if (s == "nr") nl_solver = NLSolver::NR;
      else if (s == "nrls") nl_solver = NLSolver::NRLS;
      else throw Abort("Solvers.NR.nl_solver was not provided a valid type.");
      newton_iter = (int)t.int_or("Solvers.NR.iter", 25);
      newton_rel_tol = t.num_or("Solvers.NR.rel_tol", 1e-5);
      newton_abs_tol = t.num_or("Solvers.NR.abs_tol", 1e-10);
    }
    const std::string im = lower(t.str_or("Solvers.integ_model", "FULL"));
    if (im == "full") integ_type = IntegrationType::FULL;
    else if (im == "bbar") {
      integ_type = IntegrationType::BBAR;
      if (assembly == Assembly::PA) throw Abort("Solvers.integ_model can't be BBAR if Solvers.assembly is PA.");
    } else throw Abort("Solvers.integ_model was not provided a valid type.");
    if (t.has_table("Solvers.Krylov")) {
      krylov_iter = (int)t.int_or("Solvers.Krylov.iter", 200);
      krylov_rel_tol = t.num_or("Solvers.Krylov.rel_tol", 1e-10);
      krylov_abs_tol = t.num_or("Solvers.Krylov.abs_tol", 1e-30);
      const std::string s = lower(t.str_or("Solvers.Krylov.solver", "GMRES"));
      if (s == "gmres") solver = KrylovSolver::GMRES;
      else if (s == "pcg") solver = KrylovSolver::PCG;
      else if (s == "minres") solver = KrylovSolver::MINRES;
      else throw Abort("Solvers.Krylov.solver was not provided a valid type.");
    }
  }

  void get_mesh(const Toml& t) {
    if (!t.has_table("Mesh")) throw Abort("Mesh table was not provided in toml file");
    ser_ref_levels = (int)t.int_or("Mesh.ref_ser", 0);
    par_ref_levels = (int)t.int_or("Mesh.ref_par", 0);
    order = (int)t.int_or("Mesh.p_refinement", 1);
    mesh_file = t.str_or("Mesh.floc", "../../data/cube-hex-ro.mesh");
    mesh_type = lower(t.str_or("Mesh.type", "other"));
    if (mesh_type == "auto") {
      if (!t.has_table("Mesh.Auto")) throw Abort("Mesh.type was set to Auto but Mesh.Auto does not exist");
      std::vector<double> l, n;
      if (const TVal* v = t.get("Mesh.Auto.length")) Toml::flatten(*v, l);
      if (l.size() != 3) throw Abort("Mesh.Auto.length was not provided a valid array of size 3.");
      if (const TVal* v = t.get("Mesh.Auto.ncuts")) Toml::flatten(*v, n);
      if (n.size() != 3) throw Abort("Mesh.Auto.ncuts was not provided a valid array of size 3.");
      for (int i = 0; i < 3; ++i) { mxyz[i] = l[i]; nxyz[i] = (int)n[i]; }
    } else if (mesh_type == "cubit" || mesh_type == "other") {
      throw Abort("Mesh.type = " + mesh_type + ": mesh files need MFEM's readers; this driver generates the Mesh.type = \"auto\" voxel mesh only");
    } else throw Abort("Mesh.type was not provided a valid type.");
    if (order != 1) throw Abort("Mesh.p_refinement: the kernels are written for p = 1 hexahedra");
  }
};

// whitespace-separated numbers (props / state / orientation / grain-map / custom-dt files)
inline std::vector<double> load_numbers(const std::string& path, long count, const char* what) {
  std::ifstream f(path);
  if (!f) throw Abort(std::string("Cannot open ") + what + " file: " + path);
  std::vector<double> out;
  double v;
  while ((count < 0 || (long)out.size() < count) && (f >> v)) out.push_back(v);
  if (count >= 0 && (long)out.size() != count)
    throw Abort(std::string(what) + " file " + path + " holds " + std::to_string(out.size()) + " values, " + std::to_string(count) + " expected");
  return out;
}

}  // namespace exaopt
