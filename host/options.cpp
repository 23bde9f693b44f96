#include "options.hpp"

#include <charconv>
#include <cstdlib>
#include <cstring>
#include <format>

#include "errors.hpp"

namespace lowcut {
namespace {

struct Spec {
	const char* long_name;
	char short_name;
	bool takes_value;
	const char* help;
};

constexpr Spec SPECS[] = {
	{"frequency", 'f', true, "Filter cutoff frequency in Hz. (=15)"},
	{"slope", 's', true, "Filter slope width in Hz. (=10)"},
	{"normalize", 'n', false, "Normalize output to maximum level."},
	{"verbose", 'v', false, "Verbose output."},
	{"threads", 't', true, "Accepted for compatibility; the FIR runs on the GPU grid. (=0)"},
	{"overwrite", 'O', false, "Overwrite existing files."},
	{"gpus", 'g', true, "Number of B200 GPUs to use (0 = all). (=0)"},
	{"help", 'h', false, "Display this help message."},
};

const Spec* find_long(std::string_view name)
{
	const Spec* hit = nullptr;
	for (const Spec& s : SPECS) {
		if (name == s.long_name) return &s;
		if (std::string_view(s.long_name).starts_with(name) && !name.empty()) {
			if (hit) throw UsageError(std::format("option '--{}' is ambiguous", name));
			hit = &s;
		}
	}
	return hit;
}

const Spec* find_short(char c)
{
	for (const Spec& s : SPECS)
		if (s.short_name == c) return &s;
	return nullptr;
}

double to_double(const Spec& s, const std::string& v)
{
	char* end = nullptr;
	const double d = std::strtod(v.c_str(), &end);
	if (v.empty() || end != v.c_str() + v.size())
		throw UsageError(std::format("the argument ('{}') for option '--{}' is invalid", v, s.long_name));
	return d;
}

unsigned to_unsigned(const Spec& s, const std::string& v)
{
	unsigned u = 0;
	auto [p, ec] = std::from_chars(v.data(), v.data() + v.size(), u);
	if (ec != std::errc() || p != v.data() + v.size() || v.empty())
		throw UsageError(std::format("the argument ('{}') for option '--{}' is invalid", v, s.long_name));
	return u;
}

void apply(CliOptions& o, const Spec& s, const std::string& v)
{
	switch (s.short_name) {
	case 'f': o.freq = to_double(s, v); break;
	case 's': o.slope = to_double(s, v); break;
	case 'n': o.normalize = true; break;
	case 'v': o.verbose = true; break;
	case 't': o.num_threads = to_unsigned(s, v); break;
	case 'O': o.overwrite = true; break;
	case 'g': o.gpus = to_unsigned(s, v); break;
	case 'h': o.help = true; break;
	}
}

} // namespace

CliOptions parse_cli(int argc, char** argv)
{
	CliOptions o;
	bool only_paths = false;
	for (int i = 1; i < argc; ++i) {
		const std::string a = argv[i];
		if (only_paths || a.size() < 2 || a[0] != '-') {
			o.paths.push_back(a); // a lone "-" is a path too
			continue;
		}
		if (a == "--") {
			only_paths = true;
			continue;
		}
		if (a[1] == '-') {
			const auto eq = a.find('=');
			const std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
			const Spec* s = find_long(name);
			if (!s) throw UsageError(std::format("unrecognised option '--{}'", name));
			if (!s->takes_value) {
				if (eq != std::string::npos)
					throw UsageError(std::format("option '--{}' does not take any arguments", s->long_name));
				apply(o, *s, "");
			} else if (eq != std::string::npos) {
				apply(o, *s, a.substr(eq + 1));
			} else {
				if (i + 1 >= argc)
					throw UsageError(std::format("the required argument for option '--{}' is missing", s->long_name));
				apply(o, *s, argv[++i]);
			}
			continue;
		}
		// short options, possibly bundled; a value option consumes the rest of the token
		for (size_t p = 1; p < a.size(); ++p) {
			const Spec* s = find_short(a[p]);
			if (!s) {
				// "-5" style negative numbers never occur as paths here; report like boost does
				throw UsageError(std::format("unrecognised option '-{}'", a[p]));
			}
			if (!s->takes_value) {
				apply(o, *s, "");
				continue;
			}
			std::string v = a.substr(p + 1);
			if (!v.empty() && v[0] == '=') v.erase(0, 1);
			if (v.empty()) {
				if (i + 1 >= argc)
					throw UsageError(std::format("the required argument for option '--{}' is missing", s->long_name));
				v = argv[++i];
			}
			apply(o, *s, v);
			break;
		}
	}
	return o;
}

std::string help_text()
{
	std::string t =
		"\nApplies low-cut (high-pass) FIR filter to WAVE or AIFF file.\n"
		"Usage:\n"
		"  lowcut [options] <input_file> <output_file>\n"
		"  lowcut [options] <input_file1> [input_file2 ...] <output_directory>\n"
		"Options:\n";
	for (const Spec& s : SPECS) {
		std::string left = std::format("  -{} [ --{} ]{}", s.short_name, s.long_name, s.takes_value ? " arg" : "");
		t += std::format("{:<28}{}\n", left, s.help);
	}
	return t;
}

} // namespace lowcut
