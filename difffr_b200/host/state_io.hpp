// Fluid state files next to the path: the reference's partio .bgeo particle files
// (SimulatorBase::writeFluidParticlesState / readFluidParticlesState, Simulator/SimulatorBase.cpp:2476-2606; Houdini
// "Bgeo" version 5 as read by extern/partio/src/lib/io/BGEO.cpp:173-260: big-endian header, point-attribute table,
// then per point the homogeneous position (4 floats) followed by its attributes, 32 bits each).  Only what the state
// loader consumes is kept: P, velocity, kappa, kappa_v (float -> double, row order = particle order).
#pragma once
#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace dfrhost {

struct FluidStateFile {
  int64_t n = 0;
  std::vector<double> x, v, kappa, kappa_v;  // empty when the attribute is absent
  std::vector<int32_t> id;
};

namespace bgeo_detail {
inline uint32_t be32(const unsigned char *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }
inline uint16_t be16(const unsigned char *p) { return (uint16_t)(((uint16_t)p[0] << 8) | (uint16_t)p[1]); }
inline float bef(const unsigned char *p) {
  const uint32_t u = be32(p);
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
inline void put32(std::vector<unsigned char> &o, uint32_t u) {
  o.push_back((unsigned char)(u >> 24)); o.push_back((unsigned char)(u >> 16)); o.push_back((unsigned char)(u >> 8)); o.push_back((unsigned char)u);
}
inline void put16(std::vector<unsigned char> &o, uint16_t u) { o.push_back((unsigned char)(u >> 8)); o.push_back((unsigned char)u); }
inline void putf(std::vector<unsigned char> &o, float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  put32(o, u);
}
}  // namespace bgeo_detail

inline FluidStateFile read_bgeo(const std::string &path) {
  using namespace bgeo_detail;
  std::ifstream in(path, std::ios::binary | std::ios::ate);
  if (!in) throw std::runtime_error("cannot open " + path);
  const size_t bytes = (size_t)in.tellg();
  in.seekg(0);
  std::vector<unsigned char> d(bytes);
  in.read((char *)d.data(), (std::streamsize)bytes);
  if (bytes < 41 || std::memcmp(d.data(), "Bgeo", 4) != 0 || d[4] != 'V' || be32(&d[5]) != 5)
    throw std::runtime_error(path + ": not an uncompressed Bgeo V5 file (gzip-compressed bgeo is not supported)");
  size_t p = 9;
  auto need = [&](size_t k) {
    if (p + k > bytes) throw std::runtime_error(path + ": truncated");
  };
  need(32);
  const uint32_t nPoints = be32(&d[p]);
  const uint32_t nPointAttrib = be32(&d[p + 16]);
  p += 32;
  struct Attr {
    std::string name;
    int size, type, offset;
  };
  std::vector<Attr> attrs;
  int words = 0;
  for (uint32_t a = 0; a < nPointAttrib; a++) {
    need(2);
    const uint16_t nl = be16(&d[p]);
    p += 2;
    need((size_t)nl + 6);
    Attr at;
    at.name.assign((const char *)&d[p], nl);
    p += nl;
    at.size = be16(&d[p]);
    at.type = (int)be32(&d[p + 2]);
    p += 6;
    if (at.type == 0 || at.type == 1 || at.type == 5) {
      need((size_t)at.size * 4);
      p += (size_t)at.size * 4;  // default values
    } else if (at.type == 4) {   // indexed strings: table of names follows
      need(4);
      const uint32_t ni = be32(&d[p]);
      p += 4;
      for (uint32_t k = 0; k < ni; k++) {
        need(2);
        const uint16_t l = be16(&d[p]);
        p += 2;
        need(l);
        p += l;
      }
    } else
      throw std::runtime_error(path + ": unsupported attribute type");
    at.offset = words;
    words += at.size;
    attrs.push_back(at);
  }
  const size_t stride = (size_t)(4 + words) * 4;
  need(stride * nPoints);
  FluidStateFile out;
  out.n = nPoints;
  out.x.resize(3 * (size_t)nPoints);
  const Attr *av = nullptr, *ak = nullptr, *akv = nullptr, *aid = nullptr;
  for (const Attr &a : attrs) {
    if (a.name == "velocity" && a.size == 3) av = &a;
    if (a.name == "kappa" && a.size == 1) ak = &a;
    if (a.name == "kappa_v" && a.size == 1) akv = &a;
    if (a.name == "id" && a.size == 1) aid = &a;
  }
  if (av) out.v.resize(3 * (size_t)nPoints);
  if (ak) out.kappa.resize(nPoints);
  if (akv) out.kappa_v.resize(nPoints);
  if (aid) out.id.resize(nPoints);
  for (uint32_t i = 0; i < nPoints; i++) {
    const unsigned char *r = &d[p + stride * i];
    for (int k = 0; k < 3; k++) out.x[3 * (size_t)i + k] = bef(r + 4 * k);
    const unsigned char *a0 = r + 16;
    if (av)
      for (int k = 0; k < 3; k++) out.v[3 * (size_t)i + k] = bef(a0 + 4 * (av->offset + k));
    if (ak) out.kappa[i] = bef(a0 + 4 * ak->offset);
    if (akv) out.kappa_v[i] = bef(a0 + 4 * akv->offset);
    if (aid) out.id[i] = (int32_t)be32(a0 + 4 * aid->offset);
  }
  return out;
}

// Same attribute set and order as the reference's state files (id, kappa, kappa_v, object_id, state, velocity)
inline void write_bgeo(const std::string &path, int64_t n, const double *x, const double *v, const double *kappa, const double *kappa_v) {
  using namespace bgeo_detail;
  std::vector<unsigned char> o;
  o.insert(o.end(), {'B', 'g', 'e', 'o', 'V'});
  put32(o, 5);
  put32(o, (uint32_t)n);
  put32(o, 0); put32(o, 0); put32(o, 0);  // prims, point groups, prim groups
  put32(o, 6);                            // point attributes
  put32(o, 0); put32(o, 0); put32(o, 0);  // vertex, prim, detail attributes
  struct A { const char *name; int size, type; };
  const A attrs[6] = {{"id", 1, 1}, {"kappa", 1, 0}, {"kappa_v", 1, 0}, {"object_id", 1, 1}, {"state", 1, 1}, {"velocity", 3, 5}};
  for (const A &a : attrs) {
    put16(o, (uint16_t)std::strlen(a.name));
    o.insert(o.end(), a.name, a.name + std::strlen(a.name));
    put16(o, (uint16_t)a.size);
    put32(o, (uint32_t)a.type);
    for (int k = 0; k < a.size; k++) put32(o, 0);
  }
  for (int64_t i = 0; i < n; i++) {
    for (int k = 0; k < 3; k++) putf(o, (float)x[3 * i + k]);
    putf(o, 1.0f);
    put32(o, (uint32_t)i);
    putf(o, kappa ? (float)kappa[i] : 0.f);
    putf(o, kappa_v ? (float)kappa_v[i] : 0.f);
    put32(o, 0);
    put32(o, 0);
    for (int k = 0; k < 3; k++) putf(o, v ? (float)v[3 * i + k] : 0.f);
  }
  o.push_back(0x00);  // extra-block terminators as partio writes them
  o.push_back(0xff);
  std::ofstream out(path, std::ios::binary);
  if (!out) throw std::runtime_error("cannot write " + path);
  out.write((const char *)o.data(), (std::streamsize)o.size());
}

// "<dir>/state_54.bin" -> "<dir>/state_54_particle_Fluid.bgeo" (SimulatorBase.cpp:2023-2041)
inline std::string bgeo_of_state_file(const std::string &state_file) {
  const size_t dot = state_file.find_last_of('.');
  const size_t slash = state_file.find_last_of('/');
  const std::string stem = (dot != std::string::npos && (slash == std::string::npos || dot > slash)) ? state_file.substr(0, dot) : state_file;
  return stem + "_particle_Fluid.bgeo";
}

}  // namespace dfrhost
