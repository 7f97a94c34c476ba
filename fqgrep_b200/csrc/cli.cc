// cli.cc -- `fqgrep-b200`: the reference's command line (clap `Opts`, /root/reference/src/main.rs:142-244) over the
// device path.  Only argument parsing lives here; option normalisation, the per-input loops, --count, ordered output
// and the 0 / 1 / 2 / 101 exit mapping are fqgrep_from_opts / fqgrep in include/fqgrep_b200.hpp.
//
// Accepted for compatibility and without effect on the device path: -t/--threads (the reference's worker-pool size),
// --no-order (output is always in input order).  Added: --device, --gpus, --slot-mb (the streaming ring).  --color other than "never" is rejected: the device returns selection
// bits, not match ranges (src/lib/matcher.rs:43-125 is host-side colouring, out of scope).
#include <cstdlib>
#include <cstring>
#include <string>

#include "fqgrep_b200.hpp"

using namespace fqgrep_b200;

namespace {

[[noreturn]] void usage_error(const std::string& msg) {
    std::fprintf(stderr, "error: %s\n\nUsage: fqgrep-b200 [OPTIONS] [ARGS]...\n\nFor more information, try '--help'.\n", msg.c_str());
    std::exit(2);      // clap exits 2 on a usage error
}

void help() {
    std::puts(
        "Usage: fqgrep-b200 [OPTIONS] [ARGS]...\n\n"
        "Arguments:\n"
        "  [ARGS]...  The first argument is the pattern to match, with the remaining arguments containing the files to\n"
        "             match.  If `-e` is given, then all the arguments are files to match.  Use standard input if either\n"
        "             no files are given or `-` is given.\n\n"
        "Options:\n"
        "  -t, --threads <THREADS>        accepted, ignored (matching runs on the GPU)\n"
        "      --color <COLOR>            only `never` is supported [default: never]\n"
        "  -c, --count                    Only a count of selected lines is written to standard output\n"
        "  -e, --regexp <REGEXP>          Specify a pattern used during the search of the input (repeatable)\n"
        "  -F, --fixed-strings            Interpret pattern as a set of fixed strings\n"
        "  -f, --file <FILE>              Read one or more newline separated patterns from file\n"
        "  -N, --read-names-file <FILE>   Read newline separated query names (read IDs) from file\n"
        "  -v                             Selected lines are those not matching any of the specified patterns\n"
        "  -Z, --decompress               Assume all unrecognized inputs are GZIP compressed\n"
        "      --paired                   Treat the input files as paired (R1 R2 R1 R2 ...; one file or stdin: interleaved)\n"
        "      --no-order                 accepted, ignored (output is always in input order)\n"
        "      --reverse-complement       Search the reverse complement for matches\n"
        "      --progress                 Write progress information to standard error\n"
        "      --protein                  The input files contain protein sequences\n"
        "      --iupac <IUPAC>            never | expand | regex | bit-mask [default: never]\n"
        "      --output <FILE>...         Write selected records to file(s); two files with --paired split R1 / R2\n"
        "      --device <N>               first CUDA device index [default: 0]\n"
        "      --gpus <N>                 deal the batches of every input to N devices, output still in input order [default: 1]\n"
        "      --slot-mb <MB>             size of one slot of the batch ring per input stream; bounds host and device memory [default: 64]\n"
        "  -h, --help                     Print help\n"
        "  -V, --version                  Print version");
}

}  // namespace

int main(int argc, char** argv) {
    Opts o;
    bool after_dashdash = false;
    // value of option `name` at argv[i]: "--name=value", "--name value", "-nvalue" or "-n value"
    auto value = [&](int& i, const char* arg, size_t name_len, bool is_long) -> std::string {
        if (is_long) {
            if (arg[name_len] == '=') return std::string(arg + name_len + 1);
        } else if (arg[name_len] != '\0') {
            return std::string(arg + name_len + (arg[name_len] == '=' ? 1 : 0));
        }
        if (i + 1 >= argc) usage_error(std::string("a value is required for '") + std::string(arg, name_len) + "' but none was supplied");
        return std::string(argv[++i]);
    };
    auto is_long = [](const char* a, const char* name) {
        const size_t n = std::strlen(name);
        return std::strncmp(a, name, n) == 0 && (a[n] == '\0' || a[n] == '=');
    };
    for (int i = 1; i < argc; i++) {
        const char* a = argv[i];
        if (after_dashdash || a[0] != '-' || a[1] == '\0') { o.args.push_back(a); continue; }
        if (std::strcmp(a, "--") == 0) { after_dashdash = true; continue; }
        if (a[1] == '-') {
            if (is_long(a, "--threads")) { (void)value(i, a, 9, true); }
            else if (is_long(a, "--color")) {
                const std::string v = value(i, a, 7, true);
                if (v != "never") {
                    if (v != "always" && v != "auto") usage_error("invalid value '" + v + "' for '--color <COLOR>'");
                    std::fprintf(stderr, "Error: --color %s is not supported by fqgrep-b200 (selection runs on the GPU; use --color never)\n", v.c_str());
                    return 2;
                }
            }
            else if (is_long(a, "--count")) o.count = true;
            else if (is_long(a, "--regexp")) o.regexp.push_back(value(i, a, 8, true));
            else if (is_long(a, "--fixed-strings")) o.fixed_strings = true;
            else if (is_long(a, "--file")) { o.has_file = true; o.file = value(i, a, 6, true); }
            else if (is_long(a, "--read-names-file")) { o.has_read_names_file = true; o.read_names_file = value(i, a, 17, true); }
            else if (is_long(a, "--decompress")) o.decompress = true;
            else if (is_long(a, "--paired")) o.paired = true;
            else if (is_long(a, "--no-order")) {}
            else if (is_long(a, "--reverse-complement")) o.reverse_complement = true;
            else if (is_long(a, "--progress")) o.progress = true;
            else if (is_long(a, "--protein")) o.protein = true;
            else if (is_long(a, "--iupac")) {
                const std::string v = value(i, a, 7, true);
                if (v == "never") o.iupac = IupacOption::Never;
                else if (v == "expand") o.iupac = IupacOption::Expand;
                else if (v == "regex") o.iupac = IupacOption::Regex;
                else if (v == "bit-mask") o.iupac = IupacOption::BitMask;
                else usage_error("invalid value '" + v + "' for '--iupac <IUPAC>'");
            }
            else if (is_long(a, "--output")) {      // num_args 0.. : takes the following non-option words
                if (a[8] == '=') o.output.push_back(a + 9);
                else while (i + 1 < argc && (argv[i + 1][0] != '-' || argv[i + 1][1] == '\0')) o.output.push_back(argv[++i]);
            }
            else if (is_long(a, "--device")) o.device = std::atoi(value(i, a, 8, true).c_str());
            else if (is_long(a, "--gpus")) o.gpus = std::atoi(value(i, a, 6, true).c_str());
            else if (is_long(a, "--slot-mb")) o.slot_mb = (size_t)std::atoll(value(i, a, 9, true).c_str());
            else if (is_long(a, "--help")) { help(); return 0; }
            else if (is_long(a, "--version")) { std::printf("fqgrep-b200 %s\n", fqg_version()); return 0; }
            else usage_error(std::string("unexpected argument '") + a + "' found");
            continue;
        }
        // short options, possibly bundled (-cv, -Fe PAT)
        for (size_t k = 1; a[k]; k++) {
            const char c = a[k];
            if (c == 'c') o.count = true;
            else if (c == 'F') o.fixed_strings = true;
            else if (c == 'v') o.invert_match = true;
            else if (c == 'Z') o.decompress = true;
            else if (c == 'h') { help(); return 0; }
            else if (c == 'V') { std::printf("fqgrep-b200 %s\n", fqg_version()); return 0; }
            else if (c == 't' || c == 'e' || c == 'f' || c == 'N') {
                const std::string v = value(i, a + k, 1, false);
                if (c == 'e') o.regexp.push_back(v);
                else if (c == 'f') { o.has_file = true; o.file = v; }
                else if (c == 'N') { o.has_read_names_file = true; o.read_names_file = v; }
                break;      // the value consumed the rest of this word
            }
            else usage_error(std::string("unexpected argument '-") + c + "' found");
        }
    }
    // clap `conflicts_with` (src/main.rs:182, 219, 233)
    if (o.has_read_names_file && (!o.regexp.empty() || o.has_file || o.fixed_strings))
        usage_error("the argument '--read-names-file <READ_NAMES_FILE>' cannot be used with pattern options (-e, -f, -F)");
    if (o.protein && o.reverse_complement) usage_error("the argument '--protein' cannot be used with '--reverse-complement'");
    if (o.protein && o.iupac != IupacOption::Never) usage_error("the argument '--iupac <IUPAC>' cannot be used with '--protein'");
    return fqgrep(o);
}
