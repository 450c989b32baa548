"""Minimal stand-in for the three Biopython names the reference's I/O wrappers import
(TEST INFRASTRUCTURE: lets oracle/make_golden.py import the unmodified reference in a
container without Biopython).  Behaviour relied upon: record.id = first whitespace-delimited
token of the header, record.description = whole header, writer emits '>{id} {description}'
(or the description alone if it already starts with the id) and wraps sequences at 60."""
