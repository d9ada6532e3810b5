"""A `pyminimap2.main(cmd)`-shaped entry backed by the CUDA engine, for the two command shapes of the hot path.

The reference reaches its aligner through one call, `pymm2.main(cmd) -> (stdout_paf_text, stderr_text)`; on the hot
path that is (reference src/NanoRepeat/nanoRepeat_bam.py)

    :361-362   '-c -t N -x map-ont -f 0.0 <round1_ref.fasta> <core_sequences.fastq>'          round 2, once per region
    :496-497   '-x map-ont -f 0.0 -N 100 -c --eqx -t N <round3_reference.a-b.fasta> <read.fasta>'   round 3, once per read

With `sys.modules["pyminimap2"] = nanorepeat_b200.pymm2_shim` (or `NanoRepeat.nanoRepeat_bam.pymm2 = pymm2_shim`) the
reference's UNMODIFIED round1_and_round2_estimation / round3_estimation run on top of this library: same temp files in,
PAF text out, every record the exact local alignment of the C-ABI contract (include/nanorepeat_b200.h).  This is the
compatibility path -- one launch per call, text both ways; the fast path is nanorepeat_b200.install(), which replaces the
two operators themselves.  Columns the hot path never reads (qstart, qend, n_match, align_len, mapq; paf.py:39-52) are
filled with placeholders; the strand is always '+': rounds 2-3 align cores that are already oriented (:311-312) and never
look at it.  Any other command (e.g. Step 1's anchor search, :279-281) goes to `fallback` if one was set with
set_fallback(real_pyminimap2.main), else raises NotImplementedError.
"""
from . import engine

_fallback = None


def set_fallback(fn):
    """What to call for commands that are not one of the two hot-path shapes (normally the real pyminimap2.main)."""
    global _fallback
    _fallback = fn


def _read_fasta(path):
    names, seqs = [], []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line[0] == ">":
                names.append(line[1:].split()[0] if len(line) > 1 else "")
                seqs.append([])
            elif seqs:
                seqs[-1].append(line)
    return names, ["".join(s) for s in seqs]


def _read_fastq(path):
    names, seqs = [], []
    with open(path) as f:
        while True:
            head, seq, _plus, qual = f.readline(), f.readline(), f.readline(), f.readline()
            if not qual:
                break
            names.append(head.strip()[1:].split()[0])
            seqs.append(seq.strip())
    return names, seqs


def _is_hot_path_command(toks):
    # both shapes carry '-c', '-f 0.0' and a map-ont preset and end in <template fasta> <reads>; nothing else in the
    # reference passes '-f 0.0'
    return len(toks) >= 2 and "-c" in toks and "-f" in toks and toks[toks.index("-f") + 1] == "0.0" and "-a" not in toks


def main(cmd):
    toks = cmd.split()
    if not _is_hot_path_command(toks):
        if _fallback is not None:
            return _fallback(cmd)
        raise NotImplementedError("nanorepeat_b200.pymm2_shim handles the round-2 / round-3 command shapes only "
                                  "(nanoRepeat_bam.py:361, :496); set_fallback(pyminimap2.main) for the rest")
    preset = toks[toks.index("-x") + 1] if "-x" in toks else "map-ont"
    if preset != "map-ont":
        raise NotImplementedError(f"preset {preset}: the reference maps every data type to map-ont (tk.py:502-517)")
    sc = engine.get_preset("ont")
    tfile, qfile = toks[-2], toks[-1]
    tnames, tseqs = _read_fasta(tfile)
    with open(qfile) as f:
        first = f.read(1)
    qnames, qseqs = _read_fastq(qfile) if first == "@" else _read_fasta(qfile)
    if not tseqs or not qseqs:
        return "", ""
    queries = [q for q in qseqs for _ in tseqs]
    targets = tseqs * len(qseqs)
    recs = engine.score_tasks(queries, targets, sc)
    lines = []
    nt = len(tseqs)
    for qi, (qn, qs) in enumerate(zip(qnames, qseqs)):
        block = recs[qi * nt:(qi + 1) * nt]
        order = sorted(range(nt), key=lambda i: -int(block["score"][i]))          # minimap2 prints the best hit first
        for rank, i in enumerate(order):
            s, ts, te = int(block["score"][i]), int(block["tstart"][i]), int(block["tend"][i])
            if s <= 0 or s < sc.min_dp_score:                                   # minimap2 -s: no line below it
                continue
            lines.append("\t".join(str(x) for x in (qn, len(qs), 0, len(qs), "+", tnames[i], len(tseqs[i]), ts, te,
                                                    te - ts, te - ts, 60, f"AS:i:{s}", "tp:A:P" if rank == 0 else "tp:A:S")))
    return ("\n".join(lines) + "\n" if lines else ""), ""
