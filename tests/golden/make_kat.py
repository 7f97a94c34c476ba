"""Writes tests/golden/kat_reference.json: the hot-path golden vectors of the reference's own tests,
transcribed by hand (the reference is Rust and cannot run here).  Provenance per block is the
file:line under /root/reference.  Nothing here reads /root/reference at run time."""
import json, os

T, F = True, False
kat = {
  "_provenance": "fulcrumgenomics/fqgrep v1.1.2 in-file tests; transcribed, not generated",
  # (rc, pattern, seq, expect) asserted for invert in {0,1}: expect, !expect
  "fixed_string": {"src": "src/lib/matcher.rs:786-819", "cases": [
    [F,"AG","AGG",T],[F,"CC","AGG",F],[T,"CC","AGG",T],[T,"TT","AGG",F],[F,"AT","ATGAT",T],
    [T,"CG","GCCG",T],[F,"AGAG","AGAGAGAG",T],[T,"TCTC","AGAGAGAG",T]]},
  "fixed_string_set": {"src": "src/lib/matcher.rs:825-860", "cases": [
    [F,["A","AGG","G"],"AGGG",T],[T,["A","AGG","G"],"TCCC",T],[F,["A","AGG","G"],"TTTT",F],[T,["T","AAA"],"CCCCC",F],
    [F,["AGG","C","TT"],"AGGTT",T],[T,["AGG","C","TT"],"GGGGG",T],[F,["AC","TT"],"TTACGTT",T],[T,["GT","AA"],"TTACGTT",T],
    [F,["GAGA","AGTT"],"GAGAGTT",T],[T,["CTCT","AACT"],"GAGAGTT",T]]},
  "regex": {"src": "src/lib/matcher.rs:866-898", "cases": [
    [F,"^A","AGG",T],[F,"^T","AGG",F],[T,"^C","AGG",T],[T,"^T","AGG",F],[F,"A.A","ATATA",T],[T,"T.G","CACACA",F]]},
  "regex_set": {"src": "src/lib/matcher.rs:904-940", "cases": [
    [F,["^A.G","C..","$T"],"AGGCTT",T],[T,["^T.C","..G","$A"],"AGGCTT",T],[F,["^A.G","G..","$T"],"CCTCA",F],
    [T,["$A","C.CC"],"CCTCA",F],[F,["^T",".GG","A.+G"],"ATCTACTACG",T],[T,["^C",".CC","C+.T"],"ATCTACTACG",T],
    [F,["^T","T.A"],"TTAATAA",T],[T,["^T","T.A"],"AATA",T],[F,["^T","T.+G"],"TAGAGTG",T],[T,["^A","A.+C"],"TAGAGTG",T]]},
  # (invert, names, read_id, expect); header is "@{read_id} description"
  "query_name": {"src": "src/lib/matcher.rs:946-990", "cases": [
    [F,["read1","read2"],"read1",T],[F,["read1","read2"],"read3",F],[F,["read1"],"read1_extra",F],
    [F,["read1_extra"],"read1",F],[T,["read1","read2"],"read1",F],[T,["read1","read2"],"read3",T]],
    "header_case": {"names": ["SRR001666.1"], "header": "SRR001666.1 071112_SLXA description", "expect": T}},
  # (pattern, seq, fwd_only, rc_enabled), FixedString, invert=0
  "rc_special": {"src": "src/lib/matcher.rs:1285-1373", "cases": [
    ["AAA","AAATTTCCC",T,T],["TTT","CCCGGGAAA",F,T],["AAA","CCCGGGTTT",F,T],["ACGT","ACGTACGTACGT",T,T],["T","AAAA",F,T]]},
  "revcomp": {"src": "src/lib/mod.rs:197-204", "cases": [["ACGT","ACGT"],["ACG","CGT"]]},
  "validate_fixed": {"src": "src/lib/matcher.rs:996-1040",
    "ok": [["AGTGTGATG",F],["QRNQRNQRN",T],["ARNDCEQGHILKMFPSTWYV",T],["BJOUXZ",T]],
    "err": [["AXGTGTGATG",F,"Fixed pattern must contain only DNA bases: A .. [X] .. GTGTGATG"],
            ["QRN1QRN",T,"Fixed pattern must contain only amino acids: QRN .. [1] .. QRN"]],
    "err_contains": [["ACŁT",F,"Fixed pattern must contain only DNA bases"]]},
  # files of reads; patterns via -f without -F => RegexSet; paired => (f0,f1),(f2,f3)
  "e2e_counts": {"src": "src/main.rs:926-968",
    "files": [["AAAA","TTTT"],["GGGG","CCCC"],["TTCT","CGCG"],["GGTT","GGCC"]],
    "cases": [[F,["AA"],1],[F,["CCCG"],0],[F,["T"],3],[F,["^A"],1],[F,["T.....G"],0],[F,["G"],4],[F,["GGCC","G..C"],1],
              [F,["Z","A.....G"],0],[F,["^T","AA"],3],[T,["AA"],1],[T,["CCCG"],0],[T,["CC"],2],[T,["^A"],1],[T,["T.....G"],0],
              [T,["G"],3],[T,["GGCC","G..C"],1],[T,["Z","A.....G"],0],[T,["^T","AA"],3]]},
  "e2e_order": {"src": "src/main.rs:973-1018",
    "files": [["AAAA","TTTT","ATAT","TATA","AAAT","TTTA"],["CCCC","GGGG","CGCG","GCGC","CCCG","GGGC"]],
    "cases": [[F,["A"],["AAAA","ATAT","TATA","AAAT","TTTA"]],
              [F,["A","G"],["AAAA","ATAT","TATA","AAAT","TTTA","GGGG","CGCG","GCGC","CCCG","GGGC"]],
              [F,["^A"],["AAAA","ATAT","AAAT"]],
              [F,["^A","^G"],["AAAA","ATAT","AAAT","GGGG","GCGC","GGGC"]],
              [T,["AAAA"],["AAAA","CCCC"]],
              [T,["A","G"],["AAAA","CCCC","TTTT","GGGG","ATAT","CGCG","TATA","GCGC","AAAT","CCCG","TTTA","GGGC"]],
              [T,["^AAAA"],["AAAA","CCCC"]],
              [T,["^AA","^GG"],["AAAA","CCCC","TTTT","GGGG","AAAT","CCCG","TTTA","GGGC"]]]},
  # (paired, invert, rc) -> count ; files [GGGG,GGGG] and [AAAA,CCCC]; -e TTTT
  "e2e_rc_invert_paired": {"src": "src/main.rs:1064-1099", "files": [["GGGG","GGGG"],["AAAA","CCCC"]], "pattern": "TTTT",
    "cases": [[F,F,F,0],[F,F,T,1],[F,T,F,4],[F,T,T,3],[T,F,F,0],[T,F,T,1],[T,T,F,2],[T,T,T,2]]},
  "e2e_split_output": {"src": "src/main.rs:1119-1162", "r1": ["GGGG","AAAA","AGGG","TGCA"], "r2": ["GGGG","GGGA","CCCC","ACGT"],
    "pattern": "GGG", "count": 3, "out_r1": ["GGGG","AAAA","AGGG"], "out_r2": ["GGGG","GGGA","CCCC"]},
  # on-disk header lines are "@@{n}" (fixture writes head "@{n}"); ids are "@0".. ; (paired, invert, names, count)
  "e2e_names": {"src": "src/main.rs:1324-1402",
    "single": {"n_reads": 4, "cases": [[F,["@0"],1],[F,["@0","@2"],2],[F,["@99"],0],[T,["@0"],3],[T,["@0","@1","@2","@3"],0]]},
    "paired": {"n_reads_per_file": 2, "cases": [[F,["@0"],1],[F,["@0","@1"],2],[T,["@0","@1"],0]]},
    "empty": {"n_reads": 2, "cases": [[F,[],0],[T,[],2]]}},
  "e2e_exit": {"src": "src/main.rs:1303-1318", "files": [["GTCAGC"],["AGTGCG"],["GGGTCTG"]],
    "cases": [{"pattern": "*", "fixed": F, "exit": 2}, {"pattern": "^T", "fixed": F, "exit": 1},
              {"pattern": "GTCAGC", "fixed": T, "exit": 0}, {"pattern": "^T", "fixed": T, "exit": 2}]},
  # (rc, pattern, seq, expect), invert=0 only
  "bitmask_iupac": {"src": "src/lib/matcher.rs:1148-1190", "cases": [
    [F,"GATK","GATG",T],[F,"GATK","GATT",T],[F,"GATK","GATA",F],[F,"GATK","GATC",F],[F,"GATR","GATGA",T],[F,"GATR","GATA",T],
    [F,"GATR","GATGT",T],[F,"N","A",T],[F,"N","C",T],[F,"N","G",T],[F,"N","T",T],[F,"ANA","ACA",T],[F,"ANA","AAA",T],[T,"GATK","CATC",T]]},
  "bitmask_set_iupac": {"src": "src/lib/matcher.rs:1234-1270", "cases": [
    [F,["GATK","ANA"],"GATG",T],[F,["GATK","ANA"],"ACA",T],[F,["GATK","ANA"],"CCCC",F],[F,["R","Y"],"A",T],[F,["R","Y"],"C",T],[F,["R","Y"],"AAAA",T]]},
  "bitmask_short_read": {"src": "src/lib/matcher.rs:1079-1110", "pattern": "ACGTACGT", "seq": "ACG", "expect": F},
  "iupac_expand": {"src": "src/lib/mod.rs:236-262", "fixed": [
      ["GATTACA", ["GATTACA"]], ["GATMACA", ["GATAACA","GATCACA"]],
      ["GRTBANA", ["GATCAAA","GATCACA","GATCAGA","GATCATA","GATGAAA","GATGACA","GATGAGA","GATGATA","GATTAAA","GATTACA","GATTAGA","GATTATA",
                   "GGTCAAA","GGTCACA","GGTCAGA","GGTCATA","GGTGAAA","GGTGACA","GGTGAGA","GGTGATA","GGTTAAA","GGTTACA","GGTTAGA","GGTTATA"]]],
    "regex": [["GATTACA","GATTACA"],["GATMACA","GAT[AC]ACA"],["GRTBANA","G[AG]T[CGT]A[ACGT]A"],["GATUCA","GATTCA"]],
    "limit": {"pattern": "NNNNNNNN", "message_contains": "10000"}},
  # every mode in {expand, regex, bit-mask} must give the same count (main.rs:1408-1522); -F set
  "e2e_iupac_modes": {"src": "src/main.rs:1408-1522", "cases": [
    {"seqs": ["GATG","GATT","GATA","GATC"], "patterns": ["GATK"], "invert": F, "count": 2},
    {"seqs": ["GATG","ACCA","TTTT","CCCC"], "patterns": ["GATK","ACS"], "invert": F, "count": 2},
    {"seqs": ["GATG","GATT","GATA","GATC"], "patterns": ["GATK"], "invert": T, "count": 2},
    {"seqs": ["AAAT","ACGT","TTTT"], "patterns": ["AN"], "invert": F, "count": 2}]},
  "e2e_protein": {"src": "src/main.rs:1023-1057", "reads": ["MQRFPW","MQRHKD","RHKDEW","WYFPQL"],
    "cases": [[T,["MQR"],["MQRFPW","MQRHKD"]],[T,["FPW"],["MQRFPW"]],[T,["ZZZ"],[]],
              [F,["^MQR"],["MQRFPW","MQRHKD"]],[F,["^MQR","^RHK"],["MQRFPW","MQRHKD","RHKDEW"]],[F,["^ZZZ","^YYY"],[]]]},
}
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_reference.json")
json.dump(kat, open(out, "w"), indent=1, ensure_ascii=True)
print("wrote", out)
