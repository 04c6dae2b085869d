class _Rec:
    def __init__(self, description, seq):
        self.description, self.seq = description, seq
        self.id = description.split()[0] if description.split() else ""


def parse(handle, fmt):
    assert fmt == "fasta"
    desc, parts = None, []
    for line in handle:
        line = line.rstrip("\r\n")
        if line.startswith(">"):
            if desc is not None:
                yield _Rec(desc, "".join(parts))
            desc, parts = line[1:].rstrip(), []     # Biopython strips trailing white space of the title line
        elif desc is not None:
            parts.append(line.strip())
    if desc is not None:
        yield _Rec(desc, "".join(parts))
