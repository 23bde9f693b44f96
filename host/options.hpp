// options.hpp -- command-line options of `lowcut`, matching the reference's CLI
// (main.cp:42-60): -f/--frequency (15), -s/--slope (10), -n/--normalize,
// -v/--verbose, -t/--threads (0), -O/--overwrite, -h/--help, positional paths.
// The reference builds this on c_lib's ProgramOptions (a boost::program_options
// wrapper); neither exists here, so the accepted syntax of boost's default style
// is re-created: `--name value`, `--name=value`, `-f 20`, `-f20`, bundled switches
// (`-nvO`), `--` ends the options, unambiguous long-option prefixes.
// One addition: -g/--gpus N (0 = every usable B200).
#pragma once
#include <string>
#include <vector>

namespace lowcut {

struct CliOptions {
	double freq = 15.0;
	double slope = 10.0;
	bool normalize = false;
	bool verbose = false;
	unsigned num_threads = 0;
	bool overwrite = false;
	bool help = false;
	unsigned gpus = 0;
	std::vector<std::string> paths;
};

// Throws UsageError on unknown options, missing or malformed values.
CliOptions parse_cli(int argc, char** argv);

// The text --help prints (banner of main.cp:26-32 plus the option table).
std::string help_text();

} // namespace lowcut
