// Minimal protobuf wire-format reader/writer (proto3, the subset spice21.proto uses).
// The reference decodes with prost 0.6 (spice21/src/proto.rs:22, spice21int/src/lib.rs:21-33); there is no protoc or
// C++ protobuf runtime in this build, so the wire format is handled by hand. Field numbers come from
// spice21/protos/{spice21,mos,bsim4}.proto.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace s21 {

struct DecodeError : std::runtime_error {
  explicit DecodeError(const std::string& m) : std::runtime_error("Decode Error: " + m) {}
};

struct PbReader {
  const uint8_t* p;
  const uint8_t* end;
  PbReader(const uint8_t* d, size_t n) : p(d), end(d + n) {}
  bool done() const { return p >= end; }
  uint64_t varint() {
    uint64_t v = 0;
    int shift = 0;
    while (true) {
      if (p >= end) throw DecodeError("truncated varint");
      uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7f) << shift;
      if (!(b & 0x80)) break;
      shift += 7;
      if (shift > 63) throw DecodeError("varint too long");
    }
    return v;
  }
  // Reads the next field header. wire: 0 varint, 1 fixed64, 2 len-delimited, 5 fixed32.
  bool next(uint32_t* field, uint32_t* wire) {
    if (done()) return false;
    uint64_t key = varint();
    *field = (uint32_t)(key >> 3);
    *wire = (uint32_t)(key & 7);
    if (*field == 0) throw DecodeError("field number 0");
    return true;
  }
  double fixed64_double() {
    if (end - p < 8) throw DecodeError("truncated fixed64");
    double d;
    std::memcpy(&d, p, 8);
    p += 8;
    return d;
  }
  PbReader sub() {
    uint64_t n = varint();
    if ((uint64_t)(end - p) < n) throw DecodeError("truncated length-delimited field");
    PbReader r(p, (size_t)n);
    p += n;
    return r;
  }
  std::string str() {
    PbReader r = sub();
    return std::string((const char*)r.p, (size_t)(r.end - r.p));
  }
  void skip(uint32_t wire) {
    switch (wire) {
      case 0: varint(); break;
      case 1: if (end - p < 8) throw DecodeError("truncated"); p += 8; break;
      case 2: sub(); break;
      case 5: if (end - p < 4) throw DecodeError("truncated"); p += 4; break;
      default: throw DecodeError("unsupported wire type");
    }
  }
  void expect(uint32_t wire, uint32_t want) {
    if (wire != want) throw DecodeError("unexpected wire type");
  }
};

// google.protobuf.DoubleValue / Int64Value / UInt64Value: { value = 1 }. An empty wrapper means 0.
inline double read_double_value(PbReader r) {
  double v = 0.0;
  uint32_t f, w;
  while (r.next(&f, &w)) {
    if (f == 1 && w == 1) v = r.fixed64_double();
    else r.skip(w);
  }
  return v;
}
inline int64_t read_int_value(PbReader r) {
  int64_t v = 0;
  uint32_t f, w;
  while (r.next(&f, &w)) {
    if (f == 1 && w == 0) v = (int64_t)r.varint();
    else r.skip(w);
  }
  return v;
}

struct PbWriter {
  std::vector<uint8_t> buf;
  void varint(uint64_t v) {
    while (v >= 0x80) { buf.push_back((uint8_t)(v | 0x80)); v >>= 7; }
    buf.push_back((uint8_t)v);
  }
  void key(uint32_t field, uint32_t wire) { varint(((uint64_t)field << 3) | wire); }
  void f_double(uint32_t field, double d) {  // proto3: zero default is still written by us only when asked
    key(field, 1);
    uint8_t b[8];
    std::memcpy(b, &d, 8);
    buf.insert(buf.end(), b, b + 8);
  }
  void f_bytes(uint32_t field, const uint8_t* d, size_t n) {
    key(field, 2);
    varint(n);
    buf.insert(buf.end(), d, d + n);
  }
  void f_string(uint32_t field, const std::string& s) { f_bytes(field, (const uint8_t*)s.data(), s.size()); }
  void f_msg(uint32_t field, const PbWriter& m) { f_bytes(field, m.buf.data(), m.buf.size()); }
  void f_packed_doubles(uint32_t field, const double* d, size_t n) {
    key(field, 2);
    varint(n * 8);
    const uint8_t* b = (const uint8_t*)d;
    buf.insert(buf.end(), b, b + n * 8);
  }
};

}  // namespace s21
