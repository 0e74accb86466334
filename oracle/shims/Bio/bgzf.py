"""Stand-in for Bio.bgzf.BgzfWriter (biopython is not installed and there is no network).

TEST INFRASTRUCTURE ONLY.  Restates the published behaviour of biopython's BgzfWriter, which
the reference uses at pop_factory.py:13,403,405,449,458: text is latin-1 encoded, buffered, and
every 65536 bytes one BGZF block is emitted (raw deflate through zlib at `compresslevel`,
wbits -15, DEF_MEM_LEVEL, strategy 0; 18-byte header with BSIZE; CRC32; ISIZE); close() flushes
the tail and appends the 28-byte EOF block.  Only the decompressed bytes enter the parity contract.
"""
import struct
import zlib

_HEAD = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00\x42\x43\x02\x00"
_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


class BgzfWriter:
    def __init__(self, filename=None, mode="w", fileobj=None, compresslevel=6):
        if filename and fileobj:
            raise ValueError("Supply either filename or fileobj, not both")
        if fileobj:
            handle = fileobj
        else:
            if "w" not in mode.lower() and "a" not in mode.lower():
                raise ValueError("Must use write or append mode, not %r" % mode)
            handle = open(filename, "ab" if "a" in mode.lower() else "wb")
        self._text = "b" not in mode.lower()
        self._handle = handle
        self._buffer = b""
        self.compresslevel = compresslevel

    def _write_block(self, block):
        assert len(block) <= 65536
        c = zlib.compressobj(self.compresslevel, zlib.DEFLATED, -15, zlib.DEF_MEM_LEVEL, 0)
        compressed = c.compress(block) + c.flush()
        del c
        if len(compressed) > 65536:
            raise RuntimeError("Didn't compress enough, try less data in this block")
        crc = zlib.crc32(block) & 0xFFFFFFFF
        bsize = struct.pack("<H", len(compressed) + 25)
        self._handle.write(_HEAD + bsize + compressed + struct.pack("<I", crc) + struct.pack("<I", len(block)))

    def write(self, data):
        if isinstance(data, str):
            data = data.encode("latin-1")
        data_len = len(data)
        if len(self._buffer) + data_len < 65536:
            self._buffer += data
        else:
            self._buffer += data
            while len(self._buffer) >= 65536:
                self._write_block(self._buffer[:65536])
                self._buffer = self._buffer[65536:]

    def flush(self):
        while len(self._buffer) >= 65536:
            self._write_block(self._buffer[:65535])
            self._buffer = self._buffer[65535:]
        self._write_block(self._buffer)
        self._buffer = b""
        self._handle.flush()

    def close(self):
        if self._buffer:
            self.flush()
        self._handle.write(_EOF)
        self._handle.flush()
        self._handle.close()

    def tell(self):
        return 0

    def seekable(self):
        return False

    def isatty(self):
        return False

    def fileno(self):
        return self._handle.fileno()

    def __enter__(self):
        return self

    def __exit__(self, type, value, traceback):
        self.close()
