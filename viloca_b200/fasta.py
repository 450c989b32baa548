"""FASTA reading/writing with the Biopython conventions VILOCA's files rely on
(SURVEY.md appendix A): record id = first whitespace-delimited token of the header,
description = the whole header line; the writer emits ``>{id} {description}`` (or the
description alone if it already starts with the id) and wraps sequences at 60 columns."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Iterator


@dataclass
class Record:
    id: str
    description: str
    seq: str


def parse(path) -> Iterator[Record]:
    """Records of a FASTA file (``path``: a file name, or an open text stream)."""
    import contextlib
    title, chunks = None, []
    with (open(path) if isinstance(path, (str, bytes)) or hasattr(path, "__fspath__") else contextlib.nullcontext(path)) as fh:
        for line in fh:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if title is not None:
                    yield _record(title, chunks)
                title, chunks = line[1:], []
            elif title is not None:
                chunks.append(line.strip())
    if title is not None:
        yield _record(title, chunks)


def _record(title: str, chunks: list) -> Record:
    first = title.split(None, 1)[0] if title.strip() else ""
    return Record(id=first, description=title, seq="".join(chunks))


def format_title(rec_id: str, description: str) -> str:
    if description and description.split(None, 1)[0] == rec_id:
        return description
    if description:
        return f"{rec_id} {description}"
    return rec_id


def format_body(seq: str, width: int = 60) -> str:
    """The sequence lines of a record (wrapped at ``width``), each with its newline."""
    return "".join(seq[i:i + width] + "\n" for i in range(0, len(seq), width))


def format_record(rec_id: str, description: str, seq: str, width: int = 60) -> str:
    return ">" + format_title(rec_id, description) + "\n" + format_body(seq, width)


def write(records: Iterable[tuple], path: str) -> int:
    """records: iterable of (id, description, sequence)."""
    n = 0
    with open(path, "w") as fh:
        for rec_id, description, seq in records:
            fh.write(format_record(rec_id, description, seq))
            n += 1
    return n
