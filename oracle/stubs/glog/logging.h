// Minimal stand-in for glog (absent from the build image) so that reference translation units
// compile UNMODIFIED into oracle/_ref.  TEST INFRASTRUCTURE ONLY.
#ifndef ORACLE_STUB_GLOG_LOGGING_H_
#define ORACLE_STUB_GLOG_LOGGING_H_
#include <cstdlib>
#include <iostream>
#include <sstream>
namespace oracle_stub {
struct NullStream {
  template <typename T> NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
struct FatalStream {
  std::ostringstream ss;
  FatalStream(const char* what, const char* file, int line) {
    ss << "CHECK failed: " << what << " at " << file << ":" << line << " ";
  }
  template <typename T> FatalStream& operator<<(const T& v) { ss << v; return *this; }
  ~FatalStream() { std::cerr << ss.str() << std::endl; std::abort(); }
};
struct Voidify { void operator&(const NullStream&) {} void operator&(const FatalStream&) {} };
template <typename T> T CheckNotNull(T p, const char* what, const char* file, int line) {
  if (p == nullptr) { FatalStream(what, file, line) << "is null"; }
  return p;
}
}  // namespace oracle_stub
#define INFO 0
#define WARNING 1
#define ERROR 2
#define LOG(level) oracle_stub::NullStream()
#define VLOG(level) oracle_stub::NullStream()
#define CHECK(cond) \
  (cond) ? (void)0 : oracle_stub::Voidify() & oracle_stub::FatalStream(#cond, __FILE__, __LINE__)
#define CHECK_OP_(a, b, op) CHECK((a) op (b))
#define CHECK_EQ(a, b) CHECK_OP_(a, b, ==)
#define CHECK_NE(a, b) CHECK_OP_(a, b, !=)
#define CHECK_GE(a, b) CHECK_OP_(a, b, >=)
#define CHECK_GT(a, b) CHECK_OP_(a, b, >)
#define CHECK_LE(a, b) CHECK_OP_(a, b, <=)
#define CHECK_LT(a, b) CHECK_OP_(a, b, <)
#define CHECK_NOTNULL(p) oracle_stub::CheckNotNull((p), #p, __FILE__, __LINE__)
#endif
