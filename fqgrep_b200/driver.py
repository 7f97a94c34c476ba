"""Function-level mirror of the reference's driver: `Opts` -> `fqgrep_from_opts` -> `fqgrep` exit status.

This is the selection / write-back glue of /root/reference/src/main.rs:375-675 (option normalisation that decides WHICH
matcher runs, the per-input loops, `--count`, ordered output, the 0 / 1 / 2 / 101 exit mapping of :295-311) over the
device path, so the reference's own integration tests can be restated one to one (tests/test_driver.py).  It is not a
CLI: argument parsing, colouring, logging and progress stay out of scope.  All matching happens on the GPU.
"""
import sys
from dataclasses import dataclass, field
from typing import List, Optional

from . import matcher as M


@dataclass
class Opts:
    """The fields of the reference's clap `Opts` (src/main.rs:142-244) that reach the hot path."""
    args: List[str] = field(default_factory=list)          # [pattern] files...
    regexp: List[str] = field(default_factory=list)        # -e
    file: Optional[str] = None                             # -f
    read_names_file: Optional[str] = None                  # -N
    fixed_strings: bool = False                            # -F
    invert_match: bool = False                             # -v
    reverse_complement: bool = False
    paired: bool = False
    count: bool = False                                    # -c
    decompress: bool = False                               # -Z
    protein: bool = False
    iupac: str = "never"                                   # never | expand | regex | bit-mask
    output: List[str] = field(default_factory=list)        # hidden --output [a [b]]
    device: int = 0


def read_patterns(path):
    """src/main.rs:246-255: one pattern per line, lines NOT trimmed."""
    with open(path, "rb") as f:
        data = f.read().decode("utf-8")
    lines = data.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    return [ln[:-1] if ln.endswith("\r") else ln for ln in lines]


def read_query_names(path):
    """src/main.rs:258-271: one name per line, trimmed, empty lines skipped."""
    names = set()
    with open(path, "rb") as f:
        for ln in f.read().decode("utf-8").split("\n"):
            t = ln.strip()
            if t:
                names.add(t.encode())
    return names


def fqgrep_from_opts(opts, stdout=None):
    """src/main.rs:375-675.  Returns the number of selected records (single-end) or pairs.  Raises FqgrepError for
    what the reference reports as `Err` (exit 2) or as a panic (exit 101, see FqgrepError.exit_code)."""
    regexp = list(opts.regexp)
    query_names = None
    if opts.read_names_file is not None:
        query_names = read_query_names(opts.read_names_file)
    elif opts.file is not None:
        regexp += read_patterns(opts.file)

    if query_names is not None or regexp:
        pattern, files = None, list(opts.args)
    else:
        if not opts.args:
            raise M.FqgrepError(-2, "Pattern must be given with -e or as the first positional argument ")
        pattern, files = opts.args[0], list(opts.args[1:])

    fixed_strings = opts.fixed_strings
    pattern, regexp, fixed_strings, bitencs = M.expand_iupac(pattern, regexp, fixed_strings, opts.iupac)

    if opts.paired and not (len(files) <= 1 or len(files) % 2 == 0):
        raise M.FqgrepError(-2, "Input files must be a multiple of two, or either a single file or standard input (assume interleaved) with --paired")
    if len(opts.output) > 2:
        raise M.FqgrepError(-2, "Expected at most 2 output files, got %d" % len(opts.output))
    if len(opts.output) == 2 and not opts.paired:
        raise M.FqgrepError(-2, "Two output files may only be given with --paired")

    if query_names is None and fixed_strings and bitencs is None:
        if pattern is not None:
            M.validate_fixed_pattern(pattern, opts.protein)
        elif regexp:
            for p in regexp:
                M.validate_fixed_pattern(p, opts.protein)
        else:
            raise M.FqgrepError(-2, "A pattern must be given as a positional argument or with -e/--regexp")

    match_opts = M.MatcherOpts(invert_match=opts.invert_match, reverse_complement=opts.reverse_complement)
    w1 = w2 = None
    close = []
    if not opts.count:
        if len(opts.output) == 2:
            w1, w2 = open(opts.output[0], "wb"), open(opts.output[1], "wb")
            close = [w1, w2]
        elif len(opts.output) == 1:
            w1 = open(opts.output[0], "wb")
            close = [w1]
        else:
            w1 = stdout if stdout is not None else sys.stdout.buffer
    try:
        if query_names is not None:
            matcher = M.MatcherFactory.new_query_name_matcher(query_names, match_opts)
        else:
            matcher = M.MatcherFactory.new_matcher(pattern, fixed_strings, regexp, bitencs, match_opts)

        num_matches = 0
        if not files:
            files = ["-"]

        def run(r1, r2, mode):
            nonlocal num_matches
            res = matcher.grep_fastq(r1, r2, mode=mode, device=opts.device)
            num_matches += res["count"]
            if w1 is not None and res["count"]:
                if w2 is not None:
                    a, b = M.write_selected(res["mask_words"], r1, r2, mode=mode, split=True)
                    w1.write(a)
                    w2.write(b)
                else:
                    w1.write(M.write_selected(res["mask_words"], r1, r2, mode=mode))

        if opts.paired:
            if len(files) == 1:
                run(M.read_input(files[0], opts.decompress, pinned=True), None, M.INTERLEAVED)
            else:
                for i in range(0, len(files) - 1, 2):
                    run(M.read_input(files[i], opts.decompress, pinned=True), M.read_input(files[i + 1], opts.decompress, pinned=True), M.PAIRED)
        else:
            for f in files:
                run(M.read_input(f, opts.decompress, pinned=True), None, M.SINGLE)
        if opts.count:
            out = stdout if stdout is not None else sys.stdout.buffer
            out.write(b"%d\n" % num_matches)
        return num_matches
    finally:
        for f in close:
            f.close()


def fqgrep(opts, stdout=None, stderr=None):
    """src/main.rs:295-311: None if something was selected, 1 if nothing was, 2 on an error, 101 where the reference panics."""
    err = stderr if stderr is not None else sys.stderr
    try:
        n = fqgrep_from_opts(opts, stdout)
    except M.FqgrepError as e:
        if e.exit_code == 101:
            err.write("Error: fqgrep panicked.  Please report this as a bug!\n")
            return 101
        err.write("Error: %s\n" % e.message)
        return 2
    except OSError as e:
        err.write("Error: %s\n" % e)
        return 2
    return 1 if n == 0 else None
