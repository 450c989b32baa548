from .Seq import Seq
from .SeqRecord import SeqRecord


def parse(handle, fmt):
    assert fmt == "fasta"
    own = isinstance(handle, str)
    fh = open(handle) if own else handle
    try:
        title, chunks = None, []
        for line in fh:
            line = line.rstrip("\n").rstrip("\r")
            if line.startswith(">"):
                if title is not None:
                    yield _record(title, chunks)
                title, chunks = line[1:], []
            elif title is not None:
                chunks.append(line.strip())
        if title is not None:
            yield _record(title, chunks)
    finally:
        if own:
            fh.close()


def _record(title, chunks):
    first = title.split(None, 1)[0] if title.strip() else ""
    return SeqRecord(Seq("".join(chunks)), id=first, name=first, description=title)


def write(records, handle, fmt):
    assert fmt == "fasta"
    own = isinstance(handle, str)
    fh = open(handle, "w") if own else handle
    n = 0
    try:
        for rec in records:
            rid, desc = rec.id, rec.description
            if desc and desc.split(None, 1)[0] == rid:
                title = desc
            elif desc:
                title = f"{rid} {desc}"
            else:
                title = rid
            fh.write(f">{title}\n")
            s = str(rec.seq)
            for i in range(0, len(s), 60):
                fh.write(s[i:i + 60] + "\n")
            n += 1
    finally:
        if own:
            fh.close()
    return n
