// ob_host.cpp -- host layer implementation (see ob_host.hpp for the reference map).
#include "ob_host.hpp"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

namespace optimet_b200 {

namespace constant {
// srcAna/constants.cpp:19-32
const t_real pi = 3.14159265358979323846;
const t_real c = 299792458;
const t_real mu0 = 4.0 * pi * 1e-7;
const t_real epsilon0 = 1.0 / (mu0 * c * c);
const t_real from_nm_to_m = 1e-9;
} // namespace constant

Spherical toSpherical(Cartesian const &p) {
  t_real r = std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z);
  if(r > 0.0)
    return Spherical(r, std::acos(p.z / r), std::atan2(p.y, p.x));
  return Spherical(0, 0, 0);
}
Cartesian toCartesian(Spherical const &p) {
  Cartesian c;
  c.x = p.rrr * std::sin(p.the) * std::cos(p.phi);
  c.y = p.rrr * std::sin(p.the) * std::sin(p.phi);
  c.z = p.rrr * std::cos(p.the);
  return c;
}

// ---------------------------------------------------------------------------------------------
// ElectroMagnetic
// ---------------------------------------------------------------------------------------------
ElectroMagnetic::ElectroMagnetic() : lambda(0) { init_r(1.0, 1.0, 1.0, 1.0, 1.0, 1.0); }

void ElectroMagnetic::init_r(t_complex epsilon_r_, t_complex mu_r_, t_complex epsilon_r_SH_, t_complex ksippp_,
                             t_complex ksiparppar_, t_complex gamma_) {
  epsilon_r = epsilon_r_;
  mu_r = mu_r_;
  epsilon = epsilon_r * constant::epsilon0;
  mu = mu_r * constant::mu0;
  epsilon_r_SH = epsilon_r_SH_;
  mu_r_SH = mu_r_;
  epsilon_SH = epsilon_r_SH * constant::epsilon0;
  mu_SH = mu_r_SH * constant::mu0;
  ksippp = ksippp_;
  ksiparppar = ksiparppar_;
  gamma = gamma_;
  modelType = 0;
}
void ElectroMagnetic::initHydrodynamicModel_r(t_complex a_, t_complex b_, t_complex d_, t_complex mu_r_) {
  a_SH = a_;
  b_SH = b_;
  d_SH = d_;
  mu_r = mu_r_;
  modelType = 3;
}
void ElectroMagnetic::initSiliconModel_r(t_complex mu_r_) {
  mu_r = mu_r_;
  modelType = 4;
}
void ElectroMagnetic::update(t_real lambda_) {
  lambda = lambda_;
  if(modelType == 3)
    populateHydrodynamicModel();
  if(modelType == 4)
    populateSiliconModel();
}

// Five-pole rational fit of gold, eps(w) = 1 + sum_i (a0_i - i w a1_i) / (b0_i - i w b1_i + (-i w)^2 b2_i),
// evaluated at w and 2w; second-order susceptibilities from the hydrodynamic (a, b, d) parameters
// (srcAna/ElectroMagnetic.cpp:72-142).
void ElectroMagnetic::populateHydrodynamicModel() {
  struct Pole {
    double a0, a1, b0, b1, b2;
  };
  static const Pole poles[5] = {
      {2.000003399882560, 0.0, 1.0, 1.326291192399820e-15, 0.0},
      {1.782388034422510e+32, 0.0, 0.0, 1.122727361975370e+14, 1.0},
      {9.571140818411450e+26, 8.034165109695690e+15, 1.398566311205070e+26, 7.280057361739550e+15, 1.0},
      {3.141025290600320e+24, 1.060027902520820e+14, 5.984581206741880e+23, 4.393809682455200e+15, 1.0},
      {5.056282927859510e+31, 2.176317566053640e+16, 1.707510287416960e+31, 3.256258123271410e+15, 1.0}};
  const double input_freq = constant::c / lambda;
  const double input_omega = 2 * constant::pi * input_freq;
  const t_complex mi(0.0, -1.0);
  t_complex sumFF, sumSH;
  for(int i = 0; i < 5; ++i) {
    const Pole &p = poles[i];
    sumFF += (p.a0 + input_omega * mi * p.a1) / (p.b0 + input_omega * mi * p.b1 + std::pow(input_omega * mi, 2) * p.b2);
    sumSH += (p.a0 + 2.0 * input_omega * mi * p.a1) /
             (p.b0 + 2.0 * input_omega * mi * p.b1 + std::pow(2.0 * input_omega * mi, 2) * p.b2);
  }
  epsilon_r = 1. + sumFF;
  epsilon_r_SH = 1. + sumSH;
  epsilon_SH = epsilon_r_SH * constant::epsilon0;
  epsilon = epsilon_r * constant::epsilon0;
  const double mele = 9.10938356e-31, charge = 1.602176e-19;
  const double w2 = std::pow(2.0 * constant::pi * input_freq, 2.0);
  ksippp = -(a_SH / 4.0) * (epsilon_r - 1.0) * (charge) / (mele * w2);
  ksiparppar = -(b_SH / 2.0) * (epsilon_r - 1.0) * (charge) / (mele * w2);
  gamma = -(d_SH / 8.0) * (epsilon_r - 1.0) * (charge) / (mele * w2);
}

// Schinke et al. silicon (n, k), 0.25 .. 1.45 um in 0.01 um steps (data of srcAna/ElectroMagnetic.cpp:147-182)
static const double kSiliconNK[121][2] = {
    {1.6370, 3.5889},     {1.7370, 3.9932},     {2.0300, 4.5958},     {2.8400, 5.1961},     {4.1850, 5.3124},
    {5.0490, 4.2900},     {5.0910, 3.6239},     {5.0850, 3.2824},     {5.1350, 3.0935},     {5.2450, 2.9573},
    {5.4230, 2.9078},     {5.9140, 2.9135},     {6.8200, 2.1403},     {6.5870, 0.9840},     {6.0250, 0.5031},
    {5.6230, 0.3263},     {5.3410, 0.2413},     {5.1100, 0.1769},     {4.9320, 0.1377},     {4.7900, 0.1120},
    {4.6730, 0.0954},     {4.5720, 0.0791},     {4.4850, 0.0702},     {4.4120, 0.0598},     {4.3490, 0.0538},
    {4.2890, 0.0485},     {4.2350, 0.0438},     {4.1870, 0.0395},     {4.1450, 0.0348},     {4.1030, 0.0299},
    {4.0730, 0.0280},     {4.0380, 0.0266},     {4.0060, 0.0237},     {3.9770, 0.0219},     {3.9540, 0.0201},
    {3.9310, 0.0185},     {3.9080, 0.0173},     {3.8880, 0.0168},     {3.8690, 0.0163},     {3.8510, 0.0147},
    {3.8350, 0.0144},     {3.8170, 0.0136},     {3.8050, 0.0128},     {3.7910, 0.0120},     {3.7760, 0.0113},
    {3.7650, 0.0106},     {3.7530, 0.0100},     {3.7410, 0.0093},     {3.7300, 0.0087},     {3.7190, 0.0082},
    {3.7120, 0.0076},     {3.7010, 0.0071},     {3.6930, 0.0066},     {3.6840, 0.0061},     {3.6770, 0.0057},
    {3.6690, 0.0053},     {3.6620, 0.0049},     {3.6550, 0.0045},     {3.6460, 0.0041},     {3.6410, 0.0038},
    {3.6360, 0.0035},     {3.6280, 0.0032},     {3.6220, 0.0029},     {3.6170, 0.0026},     {3.6130, 0.0023},
    {3.6100, 0.0021},     {3.6040, 0.0019},     {3.5980, 0.0017},     {3.5970, 0.0015},     {3.5900, 0.0013},
    {3.5840, 0.0011},     {3.5840, 9.8243e-04}, {3.5780, 8.4060e-04}, {3.5820, 7.1334e-04}, {3.5790, 5.9638e-04},
    {3.5750, 4.9020e-04}, {3.5720, 3.9616e-04}, {3.5680, 3.1437e-04}, {3.5650, 2.4048e-04}, {3.5620, 1.7959e-04},
    {3.5590, 1.3043e-04}, {3.5560, 9.2450e-05}, {3.5530, 6.7820e-05}, {3.5490, 5.2168e-05}, {3.5470, 3.9770e-05},
    {3.5450, 3.0217e-05}, {3.5420, 2.2913e-05}, {3.5400, 1.7068e-05}, {3.5370, 1.2382e-05}, {3.5340, 8.6210e-06},
    {3.5330, 5.6876e-06}, {3.5300, 3.4275e-06}, {3.5270, 1.7653e-06}, {3.5260, 5.5561e-07}, {3.5240, 2.3153e-07},
    {3.5220, 1.3904e-07}, {3.5200, 8.0863e-08}, {3.5180, 4.7940e-08}, {3.5170, 2.7132e-08}, {3.5150, 1.4318e-08},
    {3.5130, 5.8798e-09}, {3.5120, 2.3352e-09}, {3.5090, 1.2714e-09}, {3.5090, 7.5284e-10}, {3.5060, 4.4799e-10},
    {3.5050, 2.7228e-10}, {3.5030, 1.5856e-10}, {3.5020, 8.7196e-11}, {3.5010, 4.2039e-11}, {3.5000, 1.8128e-11},
    {3.4990, 1.0428e-11}, {3.4970, 6.2911e-12}, {3.4960, 3.9030e-12}, {3.4960, 2.6367e-12}, {3.4960, 1.7377e-12},
    {3.4930, 1.0428e-12}, {3.4920, 6.0422e-13}, {3.4920, 4.2895e-13}, {3.4900, 2.0381e-13}, {3.4880, 1.3785e-13},
    {3.4870, 1.0901e-13}};

// linear interpolation in the table with the reference's bracketing rule (first i with
// 0.25+0.01 i <= lambda_um <= 0.25+0.01 (i+1); srcAna/ElectroMagnetic.cpp:199-231)
static void silicon_nk(double lambda_um, double &n, double &k) {
  int i = 0;
  double s1 = 0, s2 = 0;
  for(i = 0; i < 121; ++i) {
    s1 = 0.25 + i * 0.01;
    s2 = 0.25 + (i + 1) * 0.01;
    if(lambda_um >= s1 && lambda_um <= s2)
      break;
  }
  if(i >= 120) { // the reference reads past the table here (undefined behaviour); flag instead
    n = k = std::nan("");
    return;
  }
  n = kSiliconNK[i][0] + ((kSiliconNK[i + 1][0] - kSiliconNK[i][0]) / (s2 - s1)) * (lambda_um - s1);
  k = kSiliconNK[i][1] + ((kSiliconNK[i + 1][1] - kSiliconNK[i][1]) / (s2 - s1)) * (lambda_um - s1);
}
void ElectroMagnetic::populateSiliconModel() {
  const double lambdaumFF = lambda * 1e6, lambdaumSH = lambdaumFF / 2.0;
  double nFF, kFF, nSH, kSH;
  silicon_nk(lambdaumFF, nFF, kFF);
  silicon_nk(lambdaumSH, nSH, kSH);
  epsilon_r = (std::pow(nFF, 2) - std::pow(kFF, 2)) + t_complex(0.0, 1.0) * (2.0 * nFF * kFF);
  epsilon_r_SH = (std::pow(nSH, 2) - std::pow(kSH, 2)) + t_complex(0.0, 1.0) * (2.0 * nSH * kSH);
  epsilon_SH = epsilon_r_SH * constant::epsilon0;
  epsilon = epsilon_r * constant::epsilon0;
  ksippp = 65e-19; // ElectroMagnetic.cpp:238-240
  ksiparppar = 3.5e-19;
  gamma = 1.3e-19;
}

// ---------------------------------------------------------------------------------------------
// Geometry
// ---------------------------------------------------------------------------------------------
static double findDistance(Spherical const &a, Spherical const &b) { // Tools.cpp:30-36
  Cartesian p1 = toCartesian(a), p2 = toCartesian(b);
  return std::sqrt(std::pow(p2.x - p1.x, 2.0) + std::pow(p2.y - p1.y, 2.0) + std::pow(p2.z - p1.z, 2.0));
}
void Geometry::pushObject(Scatterer const &object_) {
  for(auto const &obj : objects)
    if(findDistance(obj.vR, object_.vR) <= (object_.radius + obj.radius)) {
      std::ostringstream sstr;
      Cartesian a = toCartesian(object_.vR), b = toCartesian(obj.vR);
      sstr << "The sphere at (" << a.x << ", " << a.y << ", " << a.z << ") overlaps with the one at (" << b.x << ", "
           << b.y << ", " << b.z << "), with radii " << object_.radius << " and " << obj.radius;
      throw std::runtime_error(sstr.str());
    }
  objects.push_back(object_);
}
int Geometry::nMax() const {
  int r = 0;
  for(auto const &o : objects)
    r = std::max(r, o.nMax);
  return r;
}
int Geometry::nMaxS() const {
  int r = 0;
  for(auto const &o : objects)
    r = std::max(r, o.nMaxS);
  return r;
}
void Geometry::update(std::shared_ptr<Excitation const> incWave_) {
  for(auto &object : objects)
    object.elmag.update(incWave_->lambda());
}

// ---------------------------------------------------------------------------------------------
// Excitation
// ---------------------------------------------------------------------------------------------
Excitation::Excitation(unsigned long, const t_complex Einc_[3], bool SH_cond_, Spherical vKInc_, int nMax_,
                       t_complex bgcoeff)
    : vKInc(vKInc_), SH_cond(SH_cond_), nMax(nMax_), waveK(vKInc_.rrr * bgcoeff), bgcoef(bgcoeff) {
  for(int i = 0; i < 3; ++i)
    Einc[i] = Einc_[i];
  dataIncAp.assign(nMax * (nMax + 2), t_complex(0, 0));
  dataIncBp.assign(nMax * (nMax + 2), t_complex(0, 0));
}

// Wigner d^n_{0m}(theta) and derivative by upward recursion in n (srcAna/AuxCoefficients.cpp:216-290),
// including the reference's +1e-6 nudge of theta on the poles and the m<0 symmetry.
static void wigner_d0m(int nMax, int m_in, double theta, std::vector<double> &W, std::vector<double> &dW) {
  W.assign(nMax + 1, 0.0);
  dW.assign(nMax + 1, 0.0);
  const bool negative = m_in < 0;
  const long m = std::abs(m_in);
  double the = negative ? constant::pi - theta : theta;
  if((std::abs(theta) < 1e-10) || (std::abs(theta) - constant::pi + 1e-10 > 0.0))
    the += 1e-6;
  const double x = std::cos(the);
  double fact2m = 1.0, factm = 1.0;
  for(long i = 2; i <= 2 * m; ++i)
    fact2m *= (double)i;
  for(long i = 2; i <= m; ++i)
    factm *= (double)i;
  W[m] = std::pow(2.0, -(double)m) * (std::sqrt(fact2m) / factm) * std::pow(1.0 - x, m / 2.0) *
         std::pow(1.0 + x, m / 2.0);
  long s = m;
  if(m == 0 && nMax > 0) {
    W[1] = x * W[0];
    s = 1;
  }
  const double sn = std::sin(the);
  auto next = [&](long q) { // B.22
    return ((2 * q + 1) * x * W[q] - std::sqrt((double)(q * q - m * m)) * W[q - 1]) /
           std::sqrt((double)((q + 1) * (q + 1) - m * m));
  };
  auto deriv = [&](long q, double Wnext) { // B.26
    return (((q * std::sqrt((double)((q + 1) * (q + 1) - m * m)) * Wnext) / (2 * q + 1)) -
            (((q + 1) * std::sqrt((double)(q * q * (q * q - m * m))) * W[q - 1]) / (q * (2 * q + 1)))) /
           sn;
  };
  for(; s < nMax; ++s) {
    W[s + 1] = next(s);
    dW[s] = deriv(s, W[s + 1]);
  }
  if(nMax > 0)
    dW[nMax] = deriv(nMax, next(nMax));
  if(negative)
    for(int i = 0; i <= nMax; ++i) {
      const double c = 1.0 / std::pow(-1.0, (double)i);
      W[i] *= c;
      dW[i] *= -c;
    }
}

int Excitation::populate() {
  const int N = nMax * (nMax + 2);
  dataIncAp.assign(N, t_complex(0, 0));
  dataIncBp.assign(N, t_complex(0, 0));
  const double the = vKInc.the, phi = vKInc.phi;
  const bool on_axis = std::abs(the) < 1e-10 || (std::abs(the) - constant::pi + 1e-10) > 0.0;
  // unit vectors of the spherical basis at k-hat, for Tools::toProjection (Tools.cpp:277-288)
  const double st = std::sin(the), ct = std::cos(the), sp = std::sin(phi), cp = std::cos(phi);
  for(int m = nMax; m >= -nMax; --m) {
    std::vector<double> W, dW;
    wigner_d0m(nMax, m, the, W, dW);
    for(int n = std::max(1, std::abs(m)); n <= nMax; ++n) {
      double A = 0.0; // AuxCoefficients.cpp:62-74
      if(m != 0)
        A = on_axis ? m / ct * dW[n] : m / st * W[n];
      // C_nm = (0, iA, -dW), B_nm = (0, dW, iA) in (r, theta, phi); project on Cartesian axes
      const t_complex Cthe(0.0, A), Cphi(-dW[n], 0.0), Bthe(dW[n], 0.0), Bphi(0.0, A);
      const t_complex C[3] = {ct * cp * Cthe - sp * Cphi, ct * sp * Cthe + cp * Cphi, -st * Cthe};
      const t_complex B[3] = {ct * cp * Bthe - sp * Bphi, ct * sp * Bthe + cp * Bphi, -st * Bthe};
      t_complex cdot(0, 0), bdot(0, 0);
      for(int i = 0; i < 3; ++i) {
        cdot += std::conj(C[i]) * Einc[i];
        bdot += std::conj(B[i]) * Einc[i];
      }
      const double dn = std::sqrt((2.0 * n + 1.0) / (4.0 * constant::pi * (n * (n + 1)))); // AuxCoefficients.cpp:31-39
      const t_complex ph = std::exp(t_complex(0.0, -1.0) * (double)m * phi);
      const int p = n * (n + 1) - m - 1;
      const t_complex I(0.0, 1.0);
      dataIncAp[p] = 4 * constant::pi * std::pow(-1.0, m) * std::pow(I, n) * dn * cdot * ph;
      dataIncBp[p] = 4 * constant::pi * std::pow(-1.0, m) * std::pow(I, n - 1) * dn * bdot * ph;
    }
  }
  return 0;
}

void Excitation::updateWavelength(t_real lambda_) {
  vKInc.rrr = 2 * constant::pi / lambda_;
  waveK = vKInc.rrr * bgcoef;
  populate();
}

// ---------------------------------------------------------------------------------------------
// a small XML reader: elements, attributes, nesting, comments, several top-level elements
// (the shipped inputs have no single root; pugixml accepts that, Reader.cpp:964-965)
// ---------------------------------------------------------------------------------------------
namespace {
struct XmlNode {
  std::string name;
  std::map<std::string, std::string> attr;
  std::vector<std::unique_ptr<XmlNode>> children;
  const XmlNode *child(const char *n) const {
    for(auto const &c : children)
      if(c->name == n)
        return c.get();
    return nullptr;
  }
  std::vector<const XmlNode *> all(const char *n) const {
    std::vector<const XmlNode *> r;
    for(auto const &c : children)
      if(c->name == n)
        r.push_back(c.get());
    return r;
  }
  bool has(const char *a) const { return attr.count(a) != 0; }
  std::string value(const char *a) const {
    auto it = attr.find(a);
    return it == attr.end() ? std::string() : it->second;
  }
  double as_double(const char *a) const { return std::strtod(value(a).c_str(), nullptr); }
  int as_int(const char *a) const { return (int)std::strtol(value(a).c_str(), nullptr, 10); }
};
// null-safe accessors so that chains like node.child("a").child("b").attribute("x") read naturally
const XmlNode *ch(const XmlNode *n, const char *name) { return n ? n->child(name) : nullptr; }
double dbl(const XmlNode *n, const char *a) { return n ? n->as_double(a) : 0.0; }
std::string str(const XmlNode *n, const char *a) { return n ? n->value(a) : std::string(); }

struct XmlParser {
  const std::string &s;
  size_t i;
  explicit XmlParser(const std::string &text) : s(text), i(0) {}
  void skip_ws() {
    while(i < s.size() && std::isspace((unsigned char)s[i]))
      ++i;
  }
  bool starts(const char *t) const { return s.compare(i, std::strlen(t), t) == 0; }
  void skip_misc() {
    for(;;) {
      skip_ws();
      if(starts("<!--")) {
        size_t e = s.find("-->", i);
        if(e == std::string::npos)
          throw std::runtime_error("unterminated XML comment");
        i = e + 3;
      } else if(starts("<?")) {
        size_t e = s.find("?>", i);
        if(e == std::string::npos)
          throw std::runtime_error("unterminated XML declaration");
        i = e + 2;
      } else if(starts("<!")) {
        size_t e = s.find('>', i);
        i = e == std::string::npos ? s.size() : e + 1;
      } else
        return;
    }
  }
  std::string name() {
    size_t b = i;
    while(i < s.size() && (std::isalnum((unsigned char)s[i]) || s[i] == '_' || s[i] == '.' || s[i] == '-' || s[i] == ':'))
      ++i;
    if(i == b)
      throw std::runtime_error("XML parse error: expected a name");
    return s.substr(b, i - b);
  }
  std::unique_ptr<XmlNode> element() {
    if(s[i] != '<')
      throw std::runtime_error("XML parse error: expected '<'");
    ++i;
    std::unique_ptr<XmlNode> n(new XmlNode);
    n->name = name();
    for(;;) {
      skip_ws();
      if(i >= s.size())
        throw std::runtime_error("XML parse error: unterminated tag");
      if(s[i] == '/') {
        i += 2; // "/>"
        return n;
      }
      if(s[i] == '>') {
        ++i;
        break;
      }
      std::string a = name();
      skip_ws();
      if(s[i] != '=')
        throw std::runtime_error("XML parse error: expected '='");
      ++i;
      skip_ws();
      char q = s[i];
      if(q != '"' && q != '\'')
        throw std::runtime_error("XML parse error: expected a quoted value");
      size_t e = s.find(q, i + 1);
      if(e == std::string::npos)
        throw std::runtime_error("XML parse error: unterminated attribute value");
      n->attr[a] = s.substr(i + 1, e - i - 1);
      i = e + 1;
    }
    for(;;) { // content
      size_t lt = s.find('<', i);
      if(lt == std::string::npos)
        throw std::runtime_error("XML parse error: missing closing tag for " + n->name);
      i = lt;
      if(starts("</")) {
        size_t e = s.find('>', i);
        i = e + 1;
        return n;
      }
      if(starts("<!--") || starts("<?") || starts("<!")) {
        skip_misc();
        continue;
      }
      n->children.push_back(element());
    }
  }
  std::unique_ptr<XmlNode> document() {
    std::unique_ptr<XmlNode> root(new XmlNode);
    root->name = "#document";
    for(;;) {
      skip_misc();
      if(i >= s.size())
        break;
      root->children.push_back(element());
    }
    return root;
  }
};

// Reader.cpp:585-662
Scatterer read_scatterer(const XmlNode *node, int nMax, int nMaxS) {
  Scatterer result(nMax, nMaxS);
  if(str(node, "type") != "sphere")
    throw std::runtime_error("Only type=\"sphere\" objects are supported by the B200 path (srcAna tree)");
  const double nm = constant::from_nm_to_m;
  if(const XmlNode *c = ch(node, "cartesian")) {
    Cartesian p;
    p.x = c->as_double("x") * nm;
    p.y = c->as_double("y") * nm;
    p.z = c->as_double("z") * nm;
    result.vR = toSpherical(p);
  } else if(const XmlNode *s = ch(node, "spherical"))
    result.vR = Spherical(s->as_double("rrr") * nm, s->as_double("the"), s->as_double("phi"));
  else
    result.vR = Spherical(0, 0, 0);
  if(ch(node, "properties") && ch(node, "properties")->has("radius"))
    result.radius = ch(node, "properties")->as_double("radius") * nm;
  if(ch(node, "epsilon") || ch(node, "mu")) {
    if(str(ch(node, "mu"), "type") != "relative")
      throw std::runtime_error("The type for mu must be \"relative\"");
    const t_complex aux_mu(dbl(ch(node, "mu"), "value.real"), dbl(ch(node, "mu"), "value.imag"));
    const std::string etype = str(ch(node, "epsilon"), "type");
    auto cplx_of = [&](const char *child) {
      return t_complex(dbl(ch(node, child), "value.real"), dbl(ch(node, child), "value.imag"));
    };
    if(etype == "relative") {
      result.elmag.init_r(cplx_of("epsilon"), aux_mu, cplx_of("epsilon_SH"), cplx_of("ksippp"), cplx_of("ksiparppar"),
                          cplx_of("gamma"));
    } else if(etype == "GoldModel") {
      const XmlNode *p = ch(ch(node, "epsilon"), "parameters");
      result.elmag.init_r(0.0, aux_mu, 0.0, 0.0, 0.0, 0.0);
      result.elmag.initHydrodynamicModel_r(t_complex(dbl(p, "a.real"), dbl(p, "a.imag")),
                                           t_complex(dbl(p, "b.real"), dbl(p, "b.imag")),
                                           t_complex(dbl(p, "d.real"), dbl(p, "d.imag")), aux_mu);
    } else if(etype == "SiliconModel") {
      result.elmag.init_r(0.0, aux_mu, 0.0, 0.0, 0.0, 0.0);
      result.elmag.initSiliconModel_r(aux_mu);
    } else
      throw std::runtime_error("Unknown type for epsilon");
  }
  return result;
}

// Reader.cpp:84-96 / :108-120: applied only when the type is NOT "relative" (quirk kept)
void read_background(const XmlNode *geo, Geometry &g) {
  const XmlNode *bg = ch(geo, "background");
  if(!bg)
    return;
  if(str(bg, "type") != "relative") {
    t_complex aux_epsilon(dbl(ch(bg, "epsilon"), "value.real"), dbl(ch(bg, "epsilon"), "value.imag"));
    t_complex aux_mu(dbl(ch(bg, "mu"), "value.real"), dbl(ch(bg, "mu"), "value.imag"));
    g.bground.init_r(aux_epsilon, aux_mu, 0.0, 0.0, 0.0, 0.0);
  }
}

// Reader.cpp:55-181, 521-583
std::shared_ptr<Geometry> read_geometry(const XmlNode &doc) {
  const XmlNode *sim = doc.child("simulation");
  if(!sim)
    throw std::runtime_error("Simulation parameters not defined!");
  const bool ACA_cond = str(ch(sim, "ACA"), "compression") == "yes";
  const int nMax = ch(sim, "harmonics") ? ch(sim, "harmonics")->as_int("nmax") : 0;
  const int nMaxS = 1 * nMax; // Reader.cpp:67
  const XmlNode *geo = doc.child("geometry");
  if(!geo)
    throw std::runtime_error("Geometry not defined!");
  auto result = std::make_shared<Geometry>();
  result->ACAcompression(ACA_cond);
  if(const XmlNode *st = geo->child("structure")) {
    read_background(geo, *result);
    const std::string type = str(st, "type");
    const int No = ch(st, "properties") ? ch(st, "properties")->as_int("points") : 0;
    const double d = dbl(ch(st, "properties"), "distance") * constant::from_nm_to_m;
    std::vector<Cartesian> sites;
    if(type == "cube") {
      for(int k = 0; k < No; ++k)
        for(int j = 0; j < No; ++j)
          for(int i = 0; i < No; ++i)
            sites.push_back(Cartesian{d * double(i), d * double(j), d * double(k)});
    } else if(type == "surface") {
      for(int j = 0; j < No; ++j)
        for(int i = 0; i < No; ++i)
          sites.push_back(Cartesian{d * double(i), d * double(j), 0.0});
    } else
      throw std::runtime_error("structure type \"" + type + "\" is not supported by the B200 path");
    // the reference's enumeration quirk (Reader.cpp:169-179): object k sits on site k+1, the last one at the
    // position of the template object (the origin when it has no coordinates)
    const Scatterer scatterer = read_scatterer(st->child("object"), nMax, nMaxS);
    result->pushObject(scatterer);
    for(size_t i = 1; i < sites.size(); ++i) {
      result->objects.back().vR = toSpherical(sites[i]);
      result->pushObject(scatterer);
    }
  } else {
    for(const XmlNode *node : geo->all("object"))
      result->pushObject(read_scatterer(node, nMax, nMaxS));
    read_background(geo, *result);
  }
  if(result->objects.size() == 0)
    throw std::runtime_error("No scatterers defined in input");
  return result;
}

// Reader.cpp:782-834
std::shared_ptr<Excitation> read_excitation(const XmlNode &doc, int nMax, ElectroMagnetic const &bground) {
  const XmlNode *ext = doc.child("source");
  if(!ext)
    throw std::runtime_error("Source not defined!");
  const bool SH_cond = str(ch(ext, "SHsources"), "condition") == "yes";
  const t_complex bgcoeff = std::sqrt(bground.epsilon_r * bground.mu_r);
  const double wavelength = dbl(ch(ext, "wavelength"), "value") * 1e-9;
  Spherical vKinc(2 * constant::pi / wavelength, dbl(ch(ext, "propagation"), "theta") * constant::pi / 180.0,
                  dbl(ch(ext, "propagation"), "phi") * constant::pi / 180.0);
  const XmlNode *pol = ch(ext, "polarization");
  const t_complex Eth(dbl(pol, "Etheta.real"), dbl(pol, "Etheta.imag")), Eph(dbl(pol, "Ephi.real"), dbl(pol, "Ephi.imag"));
  // Tools::toProjection of (0, E_theta, E_phi) at k-hat
  const double st = std::sin(vKinc.the), ct = std::cos(vKinc.the), sp = std::sin(vKinc.phi), cp = std::cos(vKinc.phi);
  const t_complex zero(0.0, 0.0);
  const t_complex Einc[3] = {st * cp * zero + ct * cp * Eth - sp * Eph, st * sp * zero + ct * sp * Eth + cp * Eph,
                             ct * zero - st * Eth};
  auto result = std::make_shared<Excitation>(0, Einc, SH_cond, vKinc, nMax, bgcoeff);
  result->populate();
  return result;
}

// Reader.cpp:836-906
void read_output(const XmlNode &doc, Run &run) {
  const XmlNode *out = doc.child("output");
  if(!out)
    throw std::runtime_error("Output not defined!");
  const std::string type = str(out, "type");
  if(type == "coefficients")
    run.outputType = 2;
  if(type == "field") { // Reader.cpp:846-874 (grid in nm -> m; single-mode output is not part of the B200 path)
    run.outputType = 0;
    const XmlNode *grid = ch(out, "grid");
    const char *axes[3] = {"x", "y", "z"};
    for(int a = 0; a < 3; ++a) {
      const XmlNode *ax = ch(grid, axes[a]);
      if(!ax)
        throw std::runtime_error("field output: <grid> needs x, y and z axes");
      run.params[3 * a] = ax->as_double("min") * 1e-9;
      run.params[3 * a + 1] = ax->as_double("max") * 1e-9;
      run.params[3 * a + 2] = ax->as_double("steps");
    }
    run.projection = str(ch(out, "projection"), "spherical") == "true";
  }
  if(type == "response") {
    const XmlNode *scan = ch(out, "scan");
    if(const XmlNode *w = ch(scan, "wavelength")) {
      const double lam_start = w->as_double("initial"), lam_final = w->as_double("final");
      run.params[0] = lam_start * 1e-9;
      run.params[1] = lam_final * 1e-9;
      const int stepsize = (int)w->as_double("stepsize");
      if(stepsize == 0)
        throw std::runtime_error("scan stepsize must be a non-zero integer number of nm");
      const int steps = int(lam_final - lam_start) / stepsize; // Reader.cpp:886-891
      run.params[2] = steps + 1;
      run.outputType = 11;
    }
    if(const XmlNode *r = ch(scan, "radius")) {
      run.params[3] = r->as_double("initial") * 1e-9;
      run.params[4] = r->as_double("final") * 1e-9;
      run.params[5] = r->as_double("steps");
      run.outputType = ch(scan, "wavelength") ? 112 : 12;
    }
  }
}

// Reader.cpp:917-928
BelosParams read_parameter_list(const XmlNode &doc) {
  BelosParams b;
  const XmlNode *pl = doc.child("ParameterList");
  if(!pl)
    return b;
  b.present = true;
  bool has_solver = false;
  for(const XmlNode *p : pl->all("Parameter")) {
    const std::string name = p->value("name"), v = p->value("value");
    if(name == "Solver") {
      b.solver = v;
      has_solver = true;
    } else if(name == "Convergence Tolerance")
      b.tolerance = std::strtod(v.c_str(), nullptr);
    else if(name == "Maximum Iterations")
      b.max_iterations = std::atoi(v.c_str());
    else if(name == "Num Blocks")
      b.num_blocks = std::atoi(v.c_str());
    else if(name == "Block Size")
      b.block_size = std::atoi(v.c_str());
    else if(name == "Maximum Restarts")
      b.max_restarts = std::atoi(v.c_str());
    else if(name == "Verbosity")
      b.verbosity = std::atoi(v.c_str());
  }
  if(!has_solver)
    b.solver = "scalapack";
  return b;
}
} // namespace

// Reader.cpp:939-961
Run simulation_input_string(std::string const &xml_text) {
  XmlParser parser(xml_text);
  std::unique_ptr<XmlNode> doc = parser.document();
  Run result;
  result.geometry = read_geometry(*doc);
  result.nMax = result.geometry->nMax();
  result.nMaxS = result.geometry->nMaxS();
  ElectroMagnetic bground = result.geometry->bground;
  result.excitation = read_excitation(*doc, result.nMax, bground);
  result.geometry->update(result.excitation);
  read_output(*doc, result);
  result.belos_params = read_parameter_list(*doc);
  return result;
}
Run simulation_input(std::string const &fileName_) {
  std::ifstream f(fileName_.c_str());
  if(!f) {
    std::ostringstream msg;
    msg << "Error reading or parsing input file " << fileName_ << "!";
    throw std::runtime_error(msg.str());
  }
  std::stringstream ss;
  ss << f.rdbuf();
  try {
    return simulation_input_string(ss.str());
  } catch(std::runtime_error &e) {
    if(std::string(e.what()).find("XML parse error") != std::string::npos) {
      std::ostringstream msg;
      msg << "Error reading or parsing input file " << fileName_ << "!";
      throw std::runtime_error(msg.str());
    }
    throw;
  }
}

// ---------------------------------------------------------------------------------------------
// solver::B200Matrix
// ---------------------------------------------------------------------------------------------
// Solver selection of the reference: solver::factory (Solver.cpp:30-54) returns the Belos GMRES solver when the XML
// carries a Belos list whose "Solver" is neither "eigen" nor "scalapack"; "scalapack" is pzgesv_ (a dense direct
// solve) and "eigen" / the serial build run PreconditionedMatrix::solve (PreconditionedMatrixSolver.h:45-79):
// Gmres_Zcomp(tol 1e-6, maxit 240, 2 cycles) when <ACA compression="yes">, else S.colPivHouseholderQr().solve(Q).
ob_gmres_opts default_gmres(Run const &run) {
  ob_gmres_opts o;
  o.tol = 1e-6;
  o.max_iters = 240;
  o.restart = 0;
  o.max_restarts = 2;
  const bool aca = run.geometry && run.geometry->ACA_cond_;
  const bool belos = run.belos_params.present;
  std::string const &name = run.belos_params.solver;
  if(aca) {
    // <ACA compression="yes"> wins in every solver class: Gmres_Zcomp over the compressed operator with the
    // constants hard-wired there
    o.flavour = OB_GMRES_ZCOMP;
    if(belos && name == "scalapack") { // ScalapackSolver.cpp:57-66
      o.tol = 1e-7;
      o.max_iters = 250;
      o.max_restarts = 3;
    } else if(belos && name != "eigen") { // MatrixBelosSolver.cpp:33-41
      o.max_iters = 340;
      o.max_restarts = 1;
    } // else PreconditionedMatrixSolver.h:50-56
  } else if(belos && name != "scalapack" && name != "eigen") {
    o.flavour = OB_GMRES_BELOS; // scalapack/LinearSystemSolver.hpp:94-142 with the XML list
    o.tol = run.belos_params.tolerance;
    o.max_iters = run.belos_params.max_iterations;
    o.restart = run.belos_params.num_blocks;
    o.max_restarts = run.belos_params.max_restarts;
  } else {
    o.flavour = OB_SOLVE_DIRECT; // pzgesv_ (ScalapackSolver.cpp:54-128) / colPivHouseholderQr (PreconditionedMatrixSolver.h:58,75)
  }
  return o;
}

namespace solver {

B200Matrix::B200Matrix(std::shared_ptr<Geometry> geometry_, std::shared_ptr<Excitation const> incWave_, int device)
    : geometry(geometry_), incWave(incWave_), ctx(nullptr), tables_set(false), aca_mode(-1), aca_forced(false) {
  if(ob_create(device, &ctx) != 0)
    throw std::runtime_error(ob_last_error(nullptr));
  opts.flavour = (geometry && geometry->ACA_cond_) ? OB_GMRES_ZCOMP : OB_SOLVE_DIRECT; // PreconditionedMatrixSolver.h:55-58
  opts.tol = 1e-6;
  opts.max_iters = 240;
  opts.restart = 0;
  opts.max_restarts = 2;
  for(int i = 0; i < 5; ++i)
    last_cs[i] = 0;
  last_iters[0] = last_iters[1] = 0;
}
B200Matrix::B200Matrix(Run const &run, int device) : B200Matrix(run.geometry, run.excitation, device) {
  opts = default_gmres(run);
}
B200Matrix::~B200Matrix() { ob_destroy(ctx); }

void B200Matrix::check(int rc) const {
  if(rc != 0)
    throw std::runtime_error(ob_last_error(ctx));
}
void B200Matrix::set_communicator(const char uid[128], int rank, int world) { check(ob_comm_init(ctx, uid, rank, world)); }

size_t B200Matrix::scattering_size() const {
  const size_t n = geometry->nMax();
  return 2 * n * (n + 2) * geometry->objects.size();
}

void B200Matrix::update() {
  const size_t nobj = geometry->objects.size();
  if(nobj == 0)
    throw std::runtime_error("No scatterers defined in input");
  const int nMax = geometry->objects.front().nMax, nMaxS = geometry->objects.front().nMaxS;
  for(auto const &s : geometry->objects) { // PreconditionedMatrix.cpp:1160-1165
    if(s.nMax != nMax || s.nMaxS != nMaxS)
      throw std::runtime_error("All objects must have same number of harmonics");
    // NaN = outside the tabulated range.  The SH permittivity only enters the SH part: a fundamental-only run at
    // 250-500 nm (lambda / 2 below the table) is accepted like the reference accepts it
    const bool sh = incWave->SH_cond;
    if(s.elmag.epsilon != s.elmag.epsilon || (sh && s.elmag.epsilon_SH != s.elmag.epsilon_SH))
      throw std::runtime_error("SiliconModel: wavelength (or its half) outside the tabulated 0.25-1.45 um range");
  }
  std::vector<double> xyz(3 * nobj), radius(nobj);
  std::vector<t_complex> mat[7];
  for(int i = 0; i < 7; ++i)
    mat[i].resize(nobj);
  for(size_t j = 0; j < nobj; ++j) {
    Scatterer const &s = geometry->objects[j];
    Cartesian c = toCartesian(s.vR);
    xyz[3 * j] = c.x;
    xyz[3 * j + 1] = c.y;
    xyz[3 * j + 2] = c.z;
    radius[j] = s.radius;
    mat[0][j] = s.elmag.epsilon;
    mat[1][j] = s.elmag.mu;
    // unused without SH sources: a finite placeholder keeps the unused SH Mie factors from turning into NaN
    mat[2][j] = (s.elmag.epsilon_SH != s.elmag.epsilon_SH) ? s.elmag.epsilon : s.elmag.epsilon_SH;
    mat[3][j] = s.elmag.mu_SH;
    mat[4][j] = s.elmag.ksippp;
    mat[5][j] = s.elmag.ksiparppar;
    mat[6][j] = s.elmag.gamma;
  }
  // <ACA compression="yes"> -> the compressed operator (Scattering_matrix_ACA_FF/_SH, PreconditionedMatrixSolver.h:86-97)
  const bool want_aca = aca_mode < 0 ? geometry->ACA_cond_ : aca_mode != 0;
  if(want_aca != aca_forced) {
    check(ob_set_option(ctx, "operator", want_aca ? 2 : 1));
    aca_forced = want_aca;
  }
  check(ob_set_cluster(ctx, (int)nobj, xyz.data(), radius.data(), nMax, nMaxS));
  const double waveK[2] = {incWave->waveK.real(), incWave->waveK.imag()};
  const double eps_b[2] = {geometry->bground.epsilon.real(), geometry->bground.epsilon.imag()};
  const double mu_b[2] = {geometry->bground.mu.real(), geometry->bground.mu.imag()};
  check(ob_set_frequency(ctx, incWave->omega(), waveK, eps_b, mu_b, (const double *)mat[0].data(),
                         (const double *)mat[1].data(), (const double *)mat[2].data(), (const double *)mat[3].data(),
                         (const double *)mat[4].data(), (const double *)mat[5].data(), (const double *)mat[6].data()));
  check(ob_set_incident(ctx, (const double *)incWave->dataIncAp.data(), (const double *)incWave->dataIncBp.data()));
}

void B200Matrix::solve(Vector &X_sca_, Vector &X_int_, Vector &X_sca_SH, Vector &X_int_SH,
                       std::vector<double *> CGcoeff) const {
  const size_t N1 = scattering_size();
  const size_t nS = geometry->nMaxS();
  const size_t N2 = 2 * nS * (nS + 2) * geometry->objects.size();
  // every entry is overwritten by the device -> host copies of ob_run: size only (a caller that keeps its vectors
  // across wavelengths, like the C ABI layer does with page-locked ones, pays no allocation or clearing)
  if(X_sca_.size() != N1)
    X_sca_.assign(N1, t_complex(0, 0));
  if(X_int_.size() != N1)
    X_int_.assign(N1, t_complex(0, 0));
  const bool sh = incWave->SH_cond;
  if(sh) {
    if(X_sca_SH.size() != N2)
      X_sca_SH.assign(N2, t_complex(0, 0));
    if(X_int_SH.size() != N2)
      X_int_SH.assign(N2, t_complex(0, 0));
    if(CGcoeff.size() == 9 && !tables_set) {
      const double *t[9];
      for(int i = 0; i < 9; ++i)
        t[i] = CGcoeff[i];
      check(ob_set_cg_tables(ctx, t));
      tables_set = true;
    }
  }
  check(ob_run(ctx, &opts, sh ? 1 : 0, (double *)X_sca_.data(), (double *)X_int_.data(),
               sh ? (double *)X_sca_SH.data() : nullptr, sh ? (double *)X_int_SH.data() : nullptr, last_cs, last_iters));
}

} // namespace solver

// Simulation.cpp:604-685
std::vector<ScanLine> scan_wavelengths(Run &run, solver::B200Matrix &solver, std::string const &caseFile) {
  std::ofstream outASec_FF, outSSec_FF, outSSec_SH, outASec_SH;
  const bool write = !caseFile.empty();
  if(write) {
    outASec_FF.open((caseFile + "_AbsorptionCS_FF.dat").c_str());
    outSSec_FF.open((caseFile + "_ScatteringCS_FF.dat").c_str());
    if(run.excitation->SH_cond) {
      outSSec_SH.open((caseFile + "_ScatteringCS_SH.dat").c_str());
      outASec_SH.open((caseFile + "_AbsorptionCS_SH.dat").c_str());
    }
  }
  const double lami = run.params[0], lamf = run.params[1];
  const int steps = (int)run.params[2];
  const double lams = (lamf - lami) / (steps - 1);
  std::vector<ScanLine> lines;
  for(int i = 0; i < steps; i++) {
    const double lam = lami + i * lams;
    run.excitation->updateWavelength(lam);
    run.geometry->update(run.excitation);
    solver.update(run);
    Vector X_sca, X_int, X_sca_SH, X_int_SH;
    solver.solve(X_sca, X_int, X_sca_SH, X_int_SH);
    double cs[5];
    solver.cross_sections(cs);
    ScanLine l;
    l.lambda = lam;
    l.extinction_FF = cs[0];
    l.scattering_FF = cs[1];
    l.absorption_FF = cs[2];
    l.scattering_SH = cs[3];
    l.absorption_SH = cs[4];
    l.iters_FF = solver.iterations(1);
    l.iters_SH = solver.iterations(2);
    lines.push_back(l);
    if(write) {
      outASec_FF << lam << "\t" << l.absorption_FF << std::endl;
      outSSec_FF << lam << "\t" << l.scattering_FF << std::endl;
      if(run.excitation->SH_cond) {
        outSSec_SH << lam << "\t" << l.scattering_SH << std::endl;
        outASec_SH << lam << "\t" << l.absorption_SH << std::endl;
      }
    }
  }
  return lines;
}

// OutputGrid::getPoint (OutputGrid.cpp:132-157): regular Cartesian grid, x index fastest, every coordinate shifted by
// 1e-12 m ("Correct for 0"), returned as Tools::toSpherical
std::vector<double> grid_points(const double gp[9]) {
  const int nx = (int)gp[2], ny = (int)gp[5], nz = (int)gp[8];
  const long npts = (long)(gp[2] * gp[5] * gp[8]);
  const double ax = std::abs(gp[1] - gp[0]) / (gp[2] - 1), ay = std::abs(gp[4] - gp[3]) / (gp[5] - 1),
               az = std::abs(gp[7] - gp[6]) / (gp[8] - 1);
  (void)nz;
  std::vector<double> pts(3 * (size_t)npts);
  for(long it = 0; it < npts; ++it) {
    const int c0 = (int)(it % nx), c1 = (int)((it / nx) % ny), c2 = (int)(it / ((long)nx * ny));
    Spherical s = toSpherical(Cartesian(gp[0] + c0 * ax + 1e-12, gp[3] + c1 * ay + 1e-12, gp[6] + c2 * az + 1e-12));
    pts[3 * it] = s.rrr;
    pts[3 * it + 1] = s.the;
    pts[3 * it + 2] = s.phi;
  }
  return pts;
}

namespace solver {
void B200Matrix::fields(std::vector<double> const &pts_sph, bool sh, std::vector<t_complex> &out,
                        std::vector<int> &inner) const {
  const long npts = (long)(pts_sph.size() / 3);
  out.assign(12 * (size_t)npts, t_complex(0, 0));
  inner.assign((size_t)npts, -1);
  check(ob_fields(ctx, npts, pts_sph.data(), nullptr, nullptr, nullptr, nullptr, sh ? 1 : 0, (double *)out.data(),
                  inner.data()));
}
} // namespace solver

// Simulation::field_simulation (Simulation.cpp:319-366) + Result::setFields (Result.cpp:896-934)
FieldMap field_simulation(Run &run, solver::B200Matrix &solver, std::string const &caseFile) {
  FieldMap fm;
  fm.nx = (int)run.params[2];
  fm.ny = (int)run.params[5];
  fm.nz = (int)run.params[8];
  if(fm.nx < 1 || fm.ny < 1 || fm.nz < 1)
    throw std::runtime_error("field output: grid steps must be positive");
  solver.update(run);
  Vector X_sca, X_int, X_sca_SH, X_int_SH;
  solver.solve(X_sca, X_int, X_sca_SH, X_int_SH);
  const bool sh = run.excitation->SH_cond;
  std::vector<double> pts = grid_points(run.params);
  std::vector<t_complex> all;
  solver.fields(pts, sh, all, fm.inner);
  const size_t npts = pts.size() / 3;
  std::vector<t_complex> *dst[4] = {&fm.E_FF, &fm.H_FF, &fm.E_SH, &fm.H_SH};
  for(int f = 0; f < 4; ++f)
    dst[f]->assign(3 * npts, t_complex(0, 0));
  const Cartesian c0 = toCartesian(run.geometry->objects[0].vR);
  for(size_t i = 0; i < npts; ++i) {
    for(int f = 0; f < 4; ++f)
      for(int k = 0; k < 3; ++k)
        (*dst[f])[3 * i + k] = all[(i * 4 + f) * 3 + k];
    if(run.projection) { // Result.cpp:286-296: FF fields as spherical components about object 0; SH fields not set
      const Cartesian p = toCartesian(Spherical(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
      const Spherical rel = toSpherical(Cartesian(p.x - c0.x, p.y - c0.y, p.z - c0.z));
      const double st = std::sin(rel.the), ct = std::cos(rel.the), sp = std::sin(rel.phi), cp = std::cos(rel.phi);
      for(int f = 0; f < 2; ++f) { // Tools::fromProjection (Tools.cpp:290-301)
        t_complex *v = &(*dst[f])[3 * i];
        const t_complex x = v[0], y = v[1], z = v[2];
        v[0] = st * cp * x + st * sp * y + ct * z;
        v[1] = ct * cp * x + ct * sp * y - st * z;
        v[2] = cp * y - sp * x;
      }
      for(int f = 2; f < 4; ++f)
        for(int k = 0; k < 3; ++k)
          (*dst[f])[3 * i + k] = t_complex(0, 0);
    }
  }
  if(!caseFile.empty()) {
    write_field_file(caseFile + "_FF.field", fm, fm.E_FF, fm.H_FF);
    if(sh)
      write_field_file(caseFile + "_SH.field", fm, fm.E_SH, fm.H_SH);
  }
  return fm;
}

// The reference writes <case>_FF.h5 / <case>_SH.h5 with groups Field_E, Field_H, each holding X, Y, Z/{real, imag}
// and ABS/abs as [nx][ny][nz] doubles (Output.cpp:25-54, OutputGrid.cpp:41-92).  HDF5 is not available to this
// build, so the same fourteen datasets are written, in that order and in the same C order (z index fastest), to a raw
// little-endian file behind a one-line text header; scripts/field_to_h5.py rebuilds the reference's layout with h5py.
void write_field_file(std::string const &path, FieldMap const &fm, std::vector<t_complex> const &E,
                      std::vector<t_complex> const &H) {
  std::ofstream f(path.c_str(), std::ios::binary);
  if(!f)
    throw std::runtime_error("cannot open " + path);
  f << "OPTIMET_B200_FIELD 1 " << fm.nx << " " << fm.ny << " " << fm.nz
    << " Field_E/X/real Field_E/X/imag Field_E/Y/real Field_E/Y/imag Field_E/Z/real Field_E/Z/imag Field_E/ABS/abs"
    << " Field_H/X/real Field_H/X/imag Field_H/Y/real Field_H/Y/imag Field_H/Z/real Field_H/Z/imag Field_H/ABS/abs\n";
  const size_t npts = (size_t)fm.nx * fm.ny * fm.nz;
  std::vector<double> buf(npts);
  std::vector<t_complex> const *src[2] = {&E, &H};
  for(int g = 0; g < 2; ++g)
    for(int d = 0; d < 7; ++d) {
      for(size_t it = 0; it < npts; ++it) {
        const size_t ix = it % fm.nx, iy = (it / fm.nx) % fm.ny, iz = it / ((size_t)fm.nx * fm.ny);
        const t_complex *v = &(*src[g])[3 * it];
        double val;
        if(d == 6) // OutputGrid.cpp:174
          val = std::sqrt(std::pow(std::abs(v[0]), 2) + std::pow(std::abs(v[1]), 2) + std::pow(std::abs(v[2]), 2));
        else
          val = (d & 1) ? v[d / 2].imag() : v[d / 2].real();
        buf[(ix * fm.ny + iy) * fm.nz + iz] = val;
      }
      f.write((const char *)buf.data(), (std::streamsize)(npts * sizeof(double)));
    }
}

} // namespace optimet_b200
