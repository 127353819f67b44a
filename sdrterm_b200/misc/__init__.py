"""Ingest half of the drop-in (reference: src/misc/file_util.py, src/misc/read_file.py): header
parsing and fixed-size chunk reads.  Chunks are handed on as RAW BYTES; decode, byte order,
normalisation and IQ correction run on the device."""
