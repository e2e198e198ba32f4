// integration/scoped_gil_release.hpp -- lets other Python threads run during a
// libsbn_b200 call made from the reference's python module, without touching
// src/pylibsbn.cpp ("release the GIL").
#ifndef INTEGRATION_SCOPED_GIL_RELEASE_HPP_
#define INTEGRATION_SCOPED_GIL_RELEASE_HPP_

// The CPython entry points needed to let other Python threads run during a device call.
// Weak: this file is also linked into the reference's doctest binaries, which have no
// interpreter; inside the python module (pylibsbn.cpp, unchanged) they resolve to the
// running interpreter's.
extern "C" {
struct _ts;
int Py_IsInitialized(void) __attribute__((weak));
int PyGILState_Check(void) __attribute__((weak));
struct _ts *PyEval_SaveThread(void) __attribute__((weak));
void PyEval_RestoreThread(struct _ts *) __attribute__((weak));
}

// Releases the GIL for the duration of a libsbn_b200 call made from Python (what
// py::gil_scoped_release would do in pylibsbn.cpp, src/pylibsbn.cpp:192-263, without
// touching that file): the call only reads the flattened arrays this file owns.
class ScopedGilRelease {
 public:
  ScopedGilRelease() {
    if (Py_IsInitialized && PyGILState_Check && PyEval_SaveThread && PyEval_RestoreThread &&
        Py_IsInitialized() && PyGILState_Check()) {
      state_ = PyEval_SaveThread();
    }
  }
  ~ScopedGilRelease() {
    if (state_ != nullptr) {
      PyEval_RestoreThread(state_);
    }
  }
  ScopedGilRelease(const ScopedGilRelease &) = delete;
  ScopedGilRelease &operator=(const ScopedGilRelease &) = delete;

 private:
  struct _ts *state_ = nullptr;
};

#endif  // INTEGRATION_SCOPED_GIL_RELEASE_HPP_
