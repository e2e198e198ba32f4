// Shared host-side plumbing for libsbn_b200: error type + thread-local message.
#ifndef SBNB_COMMON_HPP_
#define SBNB_COMMON_HPP_

#include <stdexcept>
#include <string>

#include "../../include/sbn_b200.h"

namespace sbnb {

// Every failure inside the library is an Error carrying the C-ABI code; the
// extern "C" layer catches it, stores the message thread-locally and returns
// the code (the reference's Failwith throws std::runtime_error, sugar.hpp:67-78).
class Error : public std::runtime_error {
 public:
  Error(int code, const std::string& message) : std::runtime_error(message), code_(code) {}
  int code() const { return code_; }

 private:
  int code_;
};

[[noreturn]] inline void Fail(int code, const std::string& message) {
  throw Error(code, message);
}

inline void Require(bool condition, const std::string& message) {
  if (!condition) Fail(SBNB_ERR_INVALID_ARGUMENT, message);
}

void SetLastError(const std::string& message);

}  // namespace sbnb

#endif  // SBNB_COMMON_HPP_
