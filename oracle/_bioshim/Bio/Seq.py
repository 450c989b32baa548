class Seq:
    """Sequence object with the slice of Bio.Seq.Seq behaviour the reference relies on:
    str(), len(), iteration/indexing (numpy turns it into a 1-D array of characters, which is
    what scipy.spatial.distance.hamming receives in learn_error_params/preparation.py:84),
    .count(), .upper(), equality and hashing by content."""

    def __init__(self, data=""):
        self._data = str(data)

    def __str__(self):
        return self._data

    def __repr__(self):
        return f"Seq({self._data!r})"

    def __len__(self):
        return len(self._data)

    def __iter__(self):
        return iter(self._data)

    def __getitem__(self, i):
        r = self._data[i]
        return Seq(r) if isinstance(i, slice) else r

    def __eq__(self, other):
        return self._data == str(other)

    def __hash__(self):
        return hash(self._data)

    def count(self, sub):
        return self._data.count(str(sub))

    def upper(self):
        return Seq(self._data.upper())

    def lower(self):
        return Seq(self._data.lower())
