// errors.hpp -- the exception vocabulary of the lowcut CLI.
//
// The reference gets these four types from c_lib's <DiskerrorExceptions.h>, which is
// not part of its tree; main.cp shows how each is used and what it maps to:
//   StopNoError  -> message on stdout, EXIT_SUCCESS   (main.cp:65,153-156; --help)
//   FileNotFound -> stderr, EXIT_FAILURE              (main.cp:90,135)
//   UsageError   -> stderr, EXIT_FAILURE              (main.cp:94,99,119,124,150)
//   FileExists   -> stderr, EXIT_FAILURE              (main.cp:104,141)
// GpuError carries a non-zero status of the C-ABI (include/fir_gpu.h) so that the
// catch ladder and exit codes stay those of main.cp:153-164.
#pragma once
#include <stdexcept>
#include <string>

namespace lowcut {

struct StopNoError : std::runtime_error {
	using std::runtime_error::runtime_error;
};

struct FileNotFound : std::runtime_error {
	explicit FileNotFound(const std::string& path) : std::runtime_error("File not found: " + path) {}
};

struct FileExists : std::runtime_error {
	explicit FileExists(const std::string& path)
		: std::runtime_error("File exists (use -O to overwrite): " + path) {}
};

struct UsageError : std::runtime_error {
	using std::runtime_error::runtime_error;
};

struct FormatError : std::runtime_error {
	using std::runtime_error::runtime_error;
};

// A worker process of the batch scenario failed; it has already reported why on stderr.
struct BatchFailed : std::runtime_error {
	BatchFailed() : std::runtime_error("") {}
};

struct GpuError : std::runtime_error {
	int code;
	GpuError(int code_, const std::string& what) : std::runtime_error(what), code(code_) {}
};

} // namespace lowcut
