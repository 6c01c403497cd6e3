// Parameter classes of the drivers with the reference's names and fields
// (auxilliary/parameters.hh:15-277) on top of a small reader for the libconfig subset that the
// reference's .cfg files use (parameters_template.cfg, measurements_template.cfg): `name = value;`,
// `name = { ... }` / `name : { ... }` groups (nested), `[a, b, ...]` arrays, strings, booleans,
// integers, floats such as 1.E-12, and `//`, `#`, `/* */` comments.  libconfig++ itself is not
// installed in the build image.  Unlike libconfig the reader accepts an integer literal where a
// float is looked up.  Errors follow the reference convention: message +
// exit(-1) (parameters.cc:25-47).
#ifndef MGMC_HOST_CONFIG_HH
#define MGMC_HOST_CONFIG_HH
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

#include "vector.hh"

namespace cfg {

struct Setting {
  enum Kind { Group, Array, Int, Float, Bool, String } kind = Group;
  std::map<std::string, std::shared_ptr<Setting>> members;
  std::vector<std::shared_ptr<Setting>> items;
  long long i = 0;
  double f = 0.0;
  bool b = false;
  std::string s;
  const Setting &operator[](const std::string &name) const {
    auto it = members.find(name);
    if (it == members.end()) throw std::runtime_error("setting '" + name + "' not found");
    return *it->second;
  }
  const Setting &operator[](const char *name) const { return (*this)[std::string(name)]; }
  const Setting &lookup(const std::string &name) const { return (*this)[name]; }
  bool exists(const std::string &name) const { return members.count(name) > 0; }
  operator int() const { return (int)as_int(); }
  operator unsigned int() const { return (unsigned int)as_int(); }
  operator double() const {
    if (kind == Float) return f;
    if (kind == Int) return (double)i;
    throw std::runtime_error("setting is not a number");
  }
  operator bool() const {
    if (kind != Bool) throw std::runtime_error("setting is not a boolean");
    return b;
  }
  operator std::string() const {
    if (kind != String) throw std::runtime_error("setting is not a string");
    return s;
  }
  const char *c_str() const {
    if (kind != String) throw std::runtime_error("setting is not a string");
    return s.c_str();
  }
  long long as_int() const {
    if (kind != Int) throw std::runtime_error("setting is not an integer");
    return i;
  }
  int getLength() const { return (int)items.size(); }
  const Setting &operator[](int k) const { return *items.at(k); }
};

class Parser {
 public:
  explicit Parser(const std::string &text) : t(text) {}
  std::shared_ptr<Setting> parse_root() {
    auto root = std::make_shared<Setting>();
    parse_members(*root, '\0');
    return root;
  }

 private:
  std::string t;
  size_t p = 0;
  void skip() {
    for (;;) {
      while (p < t.size() && std::isspace((unsigned char)t[p])) ++p;
      if (p + 1 < t.size() && t[p] == '/' && t[p + 1] == '/') {
        while (p < t.size() && t[p] != '\n') ++p;
      } else if (p < t.size() && t[p] == '#') {
        while (p < t.size() && t[p] != '\n') ++p;
      } else if (p + 1 < t.size() && t[p] == '/' && t[p + 1] == '*') {
        p += 2;
        while (p + 1 < t.size() && !(t[p] == '*' && t[p + 1] == '/')) ++p;
        p += 2;
      } else {
        return;
      }
    }
  }
  [[noreturn]] void error(const std::string &m) { throw std::runtime_error("parse error at offset " + std::to_string(p) + ": " + m); }
  void parse_members(Setting &g, char close) {
    for (;;) {
      skip();
      if (p >= t.size()) {
        if (close) error("unexpected end of file");
        return;
      }
      if (close && t[p] == close) {
        ++p;
        return;
      }
      size_t b = p;
      while (p < t.size() && (std::isalnum((unsigned char)t[p]) || t[p] == '_' || t[p] == '-' || t[p] == '*')) ++p;
      if (p == b) error("setting name expected");
      const std::string name = t.substr(b, p - b);
      skip();
      if (p >= t.size() || (t[p] != '=' && t[p] != ':')) error("'=' or ':' expected after '" + name + "'");
      ++p;
      g.members[name] = parse_value();
      skip();
      if (p < t.size() && (t[p] == ';' || t[p] == ',')) ++p;
    }
  }
  std::shared_ptr<Setting> parse_value() {
    skip();
    auto v = std::make_shared<Setting>();
    if (p >= t.size()) error("value expected");
    const char c = t[p];
    if (c == '{') {
      ++p;
      v->kind = Setting::Group;
      parse_members(*v, '}');
    } else if (c == '[' || c == '(') {
      const char close = (c == '[') ? ']' : ')';
      ++p;
      v->kind = Setting::Array;
      for (;;) {
        skip();
        if (p < t.size() && t[p] == close) {
          ++p;
          break;
        }
        v->items.push_back(parse_value());
        skip();
        if (p < t.size() && t[p] == ',') ++p;
      }
    } else if (c == '"') {
      ++p;
      v->kind = Setting::String;
      while (p < t.size() && t[p] != '"') {
        if (t[p] == '\\' && p + 1 < t.size()) ++p;
        v->s += t[p++];
      }
      ++p;
    } else if (std::isalpha((unsigned char)c)) {
      size_t b = p;
      while (p < t.size() && std::isalpha((unsigned char)t[p])) ++p;
      std::string w = t.substr(b, p - b);
      for (auto &ch : w) ch = (char)std::tolower((unsigned char)ch);
      v->kind = Setting::Bool;
      if (w == "true") v->b = true;
      else if (w == "false") v->b = false;
      else error("unknown literal '" + w + "'");
    } else {
      size_t b = p;
      while (p < t.size() && (std::isdigit((unsigned char)t[p]) || t[p] == '+' || t[p] == '-' || t[p] == '.' || t[p] == 'e' || t[p] == 'E')) ++p;
      std::string w = t.substr(b, p - b);
      if (w.empty()) error("value expected");
      if (p < t.size() && t[p] == 'L') ++p;
      if (w.find_first_of(".eE") == std::string::npos) {
        v->kind = Setting::Int;
        v->i = std::stoll(w);
      } else {
        v->kind = Setting::Float;
        v->f = std::stod(w);
      }
    }
    return v;
  }
};

inline std::shared_ptr<Setting> read_file(const std::string &filename) {
  std::ifstream in(filename);
  if (!in) throw std::ios_base::failure("cannot open file");
  std::stringstream ss;
  ss << in.rdbuf();
  return Parser(ss.str()).parse_root();
}

}  // namespace cfg

/** Base class for parameters (auxilliary/parameters.hh:15-40) */
class Parameters {
 public:
  virtual ~Parameters() = default;
  int read_from_file(const std::string filename) {  // parameters.cc:21-49
    std::string classname = typeid(*this).name();
    std::shared_ptr<cfg::Setting> root;
    try {
      root = cfg::read_file(filename);
    } catch (const std::ios_base::failure &) {
      std::cerr << "Error in class '" << classname << "': cannot open configuration file '" << filename << "'." << std::endl;
      exit(-1);
    } catch (const std::exception &e) {
      std::cerr << "Error in class '" << classname << "': cannot parse configuration file '" << filename << "' (" << e.what() << ")." << std::endl;
      exit(-1);
    }
    try {
      parse_config(*root);
    } catch (const std::exception &e) {
      std::cerr << "Error in class '" << classname << "': cannot read configuration from file '" << filename << "' (" << e.what() << ")." << std::endl;
      exit(-1);
    }
    return EXIT_SUCCESS;
  }

 protected:
  virtual void parse_config(const cfg::Setting &root) = 0;
};

class GeneralParameters : public Parameters {  // parameters.cc:52-69
 public:
  int dim = 2;
  bool do_cholesky = false, do_ssor = false, do_multigridmc = true, save_posterior_statistics = false, measure_convergence = false;
  std::string operator_name;

 protected:
  void parse_config(const cfg::Setting &root) override {
    const cfg::Setting &g = root["general"];
    dim = g["dim"];
    do_cholesky = g.lookup("do_cholesky");
    do_ssor = g.lookup("do_ssor");
    do_multigridmc = g.lookup("do_multigridmc");
    save_posterior_statistics = g.lookup("save_posterior_statistics");
    measure_convergence = g.lookup("measure_convergence");
    operator_name = g.lookup("operator").c_str();
    if (!((operator_name == "prior") || (operator_name == "posterior"))) {
      std::cout << "ERROR: operator has to be 'prior' or 'posterior'" << std::endl;
      exit(-1);
    }
    std::cout << "  dimension = " << dim << std::endl;
    std::cout << "  operator = " << operator_name << std::endl;
  }
};

class LatticeParameters : public Parameters {  // parameters.cc:72-79
 public:
  unsigned int nx = 0, ny = 0, nz = 0;

 protected:
  void parse_config(const cfg::Setting &root) override {
    const cfg::Setting &l = root["lattice"];
    nx = l.lookup("nx");
    ny = l.lookup("ny");
    nz = l.lookup("nz");
    std::cout << "  lattice size = " << nx << " x " << ny << " x " << nz << std::endl;
  }
};

enum cholesky_t { SparseFactorisation = 0, DenseFactorisation = 1 };

class CholeskyParameters : public Parameters {  // parameters.cc:82-100
 public:
  cholesky_t factorisation = SparseFactorisation;

 protected:
  void parse_config(const cfg::Setting &root) override {
    std::string fac_str = root["cholesky"].lookup("factorisation");
    if (fac_str == "sparse") factorisation = SparseFactorisation;
    else if (fac_str == "dense") factorisation = DenseFactorisation;
    else {
      std::cout << "ERROR: Unknown Cholesky factorisation: '" << fac_str << "'" << std::endl;
      exit(-1);
    }
    std::cout << "  Cholesky factorisation = " << fac_str << std::endl;
  }
};

class SmootherParameters : public Parameters {  // parameters.cc:103-111
 public:
  double omega = 1.0;
  unsigned int nsmooth = 1;

 protected:
  void parse_config(const cfg::Setting &root) override {
    const cfg::Setting &s = root["smoother"];
    omega = s.lookup("omega");
    nsmooth = s.lookup("nsmooth");
    std::cout << "  smoother/Gibbs sampler " << std::endl;
    std::cout << "    number of smoothing steps = " << nsmooth << std::endl;
    std::cout << "    overrelaxation factor = " << omega << std::endl;
  }
};

class IterativeSolverParameters : public Parameters {  // parameters.cc:114-125
 public:
  double rtol = 1e-12, atol = 1e-15;
  unsigned int maxiter = 100;
  int verbose = 0;

 protected:
  void parse_config(const cfg::Setting &root) override {
    const cfg::Setting &s = root["iterative_solver"];
    rtol = s.lookup("rtol");
    atol = s.lookup("atol");
    maxiter = s.lookup("maxiter");
    verbose = s.lookup("verbose");
    std::cout << "  iterative solver" << std::endl;
    std::cout << "    rtol = " << rtol << std::endl;
    std::cout << "    atol = " << atol << std::endl;
    std::cout << "    maxiter = " << maxiter << std::endl;
  }
};

class MultigridParameters : public Parameters {  // parameters.cc:128-172
 public:
  unsigned int nlevel = 2;
  std::string smoother = "SSOR", coarse_solver = "Cholesky";
  unsigned int npresmooth = 1, npostsmooth = 1, ncoarsesmooth = 1;
  double omega = 1.0;
  unsigned int cycle = 1;
  double coarse_scaling = 1.0;
  int verbose = 0;

 protected:
  void parse_config(const cfg::Setting &root) override {
    const cfg::Setting &m = root["multigrid"];
    nlevel = m.lookup("nlevel");
    smoother = m.lookup("smoother").c_str();
    coarse_solver = m.lookup("coarse_solver").c_str();
    npresmooth = m.lookup("npresmooth");
    npostsmooth = m.lookup("npostsmooth");
    ncoarsesmooth = m.lookup("ncoarsesmooth");
    omega = m.lookup("omega");
    cycle = m.lookup("cycle");
    coarse_scaling = m.lookup("coarse_scaling");
    verbose = m.lookup("verbose");
    if (!(smoother == "SOR" || smoother == "SSOR")) {
      std::cout << "ERROR: multigrid smoother has to be 'SOR' or 'SSOR'" << std::endl;
      exit(-1);
    }
    if (!(coarse_solver == "SSOR" || coarse_solver == "Cholesky")) {
      std::cout << "ERROR: multigrid coarse solver has to be 'SSOR' or 'Cholesky'" << std::endl;
      exit(-1);
    }
    std::cout << "  multigrid" << std::endl;
    std::cout << "    levels = " << nlevel << std::endl;
    std::cout << "    smoother = " << smoother << std::endl;
    std::cout << "    coarse solver = " << coarse_solver << std::endl;
    std::cout << "    npresmooth = " << npresmooth << ", npostsmooth = " << npostsmooth << ", ncoarsesmooth = " << ncoarsesmooth << std::endl;
    std::cout << "    omega = " << omega << ", cycle = " << cycle << ", coarse_scaling = " << coarse_scaling << std::endl;
  }
};

class SamplingParameters : public Parameters {  // parameters.cc:175-188
 public:
  unsigned int nsamples = 0, nwarmup = 0, nsamplesconvergence = 0, nstepsconvergence = 0;

 protected:
  void parse_config(const cfg::Setting &root) override {
    const cfg::Setting &s = root["sampling"];
    nsamples = s["timeseries"].lookup("nsamples");
    nwarmup = s["timeseries"].lookup("nwarmup");
    nstepsconvergence = s["convergence"].lookup("nsteps");
    nsamplesconvergence = s["convergence"].lookup("nsamples");
    std::cout << "  sampling" << std::endl;
    std::cout << "    timeseries: nsamples = " << nsamples << ", nwarmup = " << nwarmup << std::endl;
    std::cout << "    convergence: nsteps = " << nstepsconvergence << ", nsamples = " << nsamplesconvergence << std::endl;
  }
};

class PriorParameters : public Parameters {  // parameters.cc:191-213
 public:
  std::string pde_model, correlationlength_model;

 protected:
  void parse_config(const cfg::Setting &root) override {
    const cfg::Setting &p = root["prior"];
    pde_model = p.lookup("pdemodel").c_str();
    correlationlength_model = p.lookup("correlationlengthmodel").c_str();
    std::cout << "  prior: pde model = " << pde_model << ", correlation length model = " << correlationlength_model << std::endl;
  }
};

class ConstantCorrelationLengthModelParameters : public Parameters {  // parameters.cc:216-222
 public:
  double Lambda = 1.0;

 protected:
  void parse_config(const cfg::Setting &root) override { Lambda = root["constantcorrelationlengthmodel"].lookup("Lambda"); }
};

class PeriodicCorrelationLengthModelParameters : public Parameters {  // parameters.cc:225-243
 public:
  double Lambda_min = 1.0, Lambda_max = 1.0;

 protected:
  void parse_config(const cfg::Setting &root) override {
    Lambda_min = root["periodiccorrelationlengthmodel"].lookup("Lambda_min");
    Lambda_max = root["periodiccorrelationlengthmodel"].lookup("Lambda_max");
  }
};

class MeasurementParameters : public Parameters {  // parameters.cc:246-316
 public:
  int dim = 2;
  unsigned int n = 0;
  std::vector<Eigen::VectorXd> measurement_locations;
  double radius = 0.0;
  Eigen::VectorXd mean, variance;
  double variance_scaling = 1.0;
  Eigen::VectorXd sample_location;
  bool measure_global = false;
  double variance_global = 0.0, mean_global = 0.0;

 protected:
  void parse_config(const cfg::Setting &root) override {
    const cfg::Setting &m = root["measurements"];
    radius = m.lookup("radius");
    variance_scaling = m.lookup("variance_scaling");
    measure_global = m.lookup("measure_global");
    mean_global = m.lookup("mean_global");
    variance_global = m.lookup("variance_global");
    const cfg::Setting &sl = m.lookup("sample_location");
    sample_location = Eigen::VectorXd(sl.getLength());
    for (int k = 0; k < sl.getLength(); ++k) sample_location[k] = (double)sl[k];
    const std::string filename = m.lookup("filename").c_str();
    std::cout << "  measurements: radius = " << radius << ", file = " << filename << std::endl;
    std::shared_ptr<cfg::Setting> mf;
    try {
      mf = cfg::read_file(filename);
    } catch (const std::exception &) {
      std::cerr << "Error: cannot read measurement file '" << filename << "'." << std::endl;
      exit(-1);
    }
    dim = (*mf)["dim"];
    n = (*mf)["n"];
    const cfg::Setting &loc = (*mf)["measurement_locations"];
    const cfg::Setting &mu = (*mf)["mean"];
    const cfg::Setting &var = (*mf)["variance"];
    if (loc.getLength() != (int)(n * dim) || mu.getLength() != (int)n || var.getLength() != (int)n) {
      std::cerr << "Error: inconsistent array lengths in measurement file '" << filename << "'." << std::endl;
      exit(-1);
    }
    measurement_locations.clear();
    mean = Eigen::VectorXd(n);
    variance = Eigen::VectorXd(n);
    for (unsigned int k = 0; k < n; ++k) {
      Eigen::VectorXd x(dim);
      for (int d = 0; d < dim; ++d) x[d] = (double)loc[(int)(k * dim + d)];
      measurement_locations.push_back(x);
      mean[k] = (double)mu[(int)k];
      variance[k] = (double)var[(int)k];
    }
  }
};
#endif
