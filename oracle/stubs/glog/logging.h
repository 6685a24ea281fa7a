// ORACLE build stub (test infrastructure): the vendored Ceres 2.0 headers include <glog/logging.h>, which is not
// installed here.  CHECK / DCHECK / LOG / VLOG compile to no-op streams (the checks guard programmer errors only).
#pragma once
#include <iostream>
namespace oracle_glog_stub {
struct NullStream {
  template <typename T> NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
}  // namespace oracle_glog_stub
#define ORACLE_GLOG_NULL() if (true) {} else ::oracle_glog_stub::NullStream()
#define CHECK(c) ORACLE_GLOG_NULL()
#define CHECK_EQ(a, b) ORACLE_GLOG_NULL()
#define CHECK_NE(a, b) ORACLE_GLOG_NULL()
#define CHECK_LT(a, b) ORACLE_GLOG_NULL()
#define CHECK_LE(a, b) ORACLE_GLOG_NULL()
#define CHECK_GT(a, b) ORACLE_GLOG_NULL()
#define CHECK_GE(a, b) ORACLE_GLOG_NULL()
#define CHECK_NOTNULL(p) (p)
#define DCHECK(c) ORACLE_GLOG_NULL()
#define DCHECK_EQ(a, b) ORACLE_GLOG_NULL()
#define DCHECK_NE(a, b) ORACLE_GLOG_NULL()
#define DCHECK_LT(a, b) ORACLE_GLOG_NULL()
#define DCHECK_LE(a, b) ORACLE_GLOG_NULL()
#define DCHECK_GT(a, b) ORACLE_GLOG_NULL()
#define DCHECK_GE(a, b) ORACLE_GLOG_NULL()
#define LOG(s) ORACLE_GLOG_NULL()
#define VLOG(n) ORACLE_GLOG_NULL()
#define LOG_IF(s, c) ORACLE_GLOG_NULL()
