"""Tiny FASTA/FASTQ reader with the slice of the Bio.SeqIO API that the
reference calls: parse(handle_or_path, fmt) -> records with .id, .seq,
.letter_annotations['phred_quality'] (shim part 1)."""


class _Rec(object):
    __slots__ = ("id", "seq", "letter_annotations")

    def __init__(self, rid, seq, quals=None):
        self.id = rid
        self.seq = seq
        self.letter_annotations = {} if quals is None else {"phred_quality": quals}


def _lines(src):
    if hasattr(src, "read"):
        for ln in src:
            yield ln
    else:
        with open(src, "r") as fh:
            for ln in fh:
                yield ln


def _fasta(src):
    rid, chunks = None, []
    for ln in _lines(src):
        if ln.startswith(">"):
            if rid is not None:
                yield _Rec(rid, "".join(chunks))
            toks = ln[1:].split()
            rid, chunks = (toks[0] if toks else ""), []
        elif rid is not None:
            chunks.append(ln.strip())
    if rid is not None:
        yield _Rec(rid, "".join(chunks))


def _fastq(src):
    it = iter(_lines(src))
    for head in it:
        if not head.strip():
            continue
        seq = next(it).rstrip("\n")
        next(it)
        qual = next(it).rstrip("\n")
        toks = head[1:].split()
        yield _Rec(toks[0] if toks else "", seq, [ord(c) - 33 for c in qual])


def parse(src, fmt):
    if fmt == "fasta":
        return _fasta(src)
    if fmt == "fastq":
        return _fastq(src)
    raise ValueError("shim supports fasta/fastq only")
