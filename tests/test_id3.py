"""ID3 tags of the container row.  The writer restates src/main/java/mp3/ID3Tag.java; nothing in the build
image runs Java, so it is checked by an independent reader of the two published layouts written here (ID3v1: 128 fixed bytes;
ID3v2.3: 10-byte header with a 28-bit syncsafe size, frames of id[4] size[4] flags[2] body) and by the reference's own decision
rules (when a version 2 tag is written at all, ID3Tag.java:976-989)."""
import pytest

import lamejs_b200 as M

LAME = "LAME 32bits version 3.98.4 (http://www.mp3dev.org/)"


def read_v2(b):
    assert b[:3] == b"ID3" and b[3:6] == bytes([3, 0, 0])
    assert all(x < 128 for x in b[6:10])
    size = (b[6] << 21) | (b[7] << 14) | (b[8] << 7) | b[9]
    assert len(b) == size + 10
    frames, p = [], 10
    while p + 10 <= len(b) and b[p] != 0:
        fid = b[p:p + 4].decode("ascii")
        n = int.from_bytes(b[p + 4:p + 8], "big")
        assert b[p + 8:p + 10] == b"\0\0"
        body = b[p + 10:p + 10 + n]
        assert body[0] == 0                                    # ISO-8859-1
        if fid == "COMM":
            assert body[1:4] == b"XXX" and body[4] == 0
            frames.append((fid, body[5:].decode("latin-1")))
        else:
            frames.append((fid, body[1:].decode("latin-1")))
        p += 10 + n
    assert not any(b[p:])                                      # padding
    return frames, len(b) - p


def test_v1_layout():
    t = M.id3v1_tag(title="Title", artist="Artist", album="Album", year="1999", comment="A comment", track="7", genre="Rock")
    assert len(t) == 128 and t[:3] == b"TAG"
    assert t[3:33] == b"Title".ljust(30, b"\0") and t[33:63] == b"Artist".ljust(30, b"\0") and t[63:93] == b"Album".ljust(30, b"\0")
    assert t[93:97] == b"1999" and t[97:125] == b"A comment".ljust(28, b"\0") and t[125] == 0 and t[126] == 7 and t[127] == 17
    s = M.id3v1_tag(title="T", flags=M.ID3_SPACE_V1)
    assert s[3:33] == b"T".ljust(30) and s[93:97] == b"    " and s[97:127] == b" " * 30 and s[127] == 255      # no track: 30-byte comment
    assert M.id3v1_tag() == b"" and M.id3v1_tag(title="x", flags=M.ID3_V2_ONLY) == b""
    long = M.id3v1_tag(title="x" * 40, year="12345", genre="147")
    assert long[3:33] == b"x" * 30 and long[93:97] == b"9999" and long[127] == 147
    assert M.id3v1_tag(genre="No Such Genre")[127] == 12                                                        # "Other"


def test_v2_is_written_only_when_asked_for_or_needed():
    fits = dict(title="Short", artist="A", album="B", year="2001", comment="c", track="3", genre="Jazz")
    assert M.id3v2_tag(**fits) == b""
    assert M.id3v2_tag(flags=M.ID3_V1_ONLY, title="x" * 31) == b""
    for k in ("title", "artist", "album", "comment"):
        assert M.id3v2_tag(**dict(fits, **{k: "y" * 31})) != b""
    assert M.id3v2_tag(**dict(fits, comment="z" * 29)) != b"" and M.id3v2_tag(**dict(fits, comment="z" * 29, track=None)) == b""
    assert M.id3v2_tag(**dict(fits, track="3/12")) != b"" and M.id3v2_tag(**dict(fits, track="300")) != b""
    assert M.id3v2_tag(**dict(fits, genre="Shoegaze")) != b""
    assert M.id3v2_tag(flags=M.ID3_ADD_V2) != b""


def test_v2_frames_round_trip():
    b = M.id3v2_tag(flags=M.ID3_ADD_V2, num_samples=441000, samplerate=44100, title="Täst", artist="Artist", album="Album", year="1987",
                    comment="Hello, world", track="4/9", genre="17")
    frames, pad = read_v2(b)
    assert pad == 0
    assert frames == [("TSSE", LAME), ("TIT2", "Täst"), ("TPE1", "Artist"), ("TALB", "Album"), ("TYER", "1987"), ("COMM", "Hello, world"),
                      ("TRCK", "4/9"), ("TCON", "Rock"), ("TLEN", "10000")]
    frames, pad = read_v2(M.id3v2_tag(flags=M.ID3_PAD_V2, padding=300, title="x"))
    assert frames == [("TSSE", LAME), ("TIT2", "x")] and pad == 300
    frames, pad = read_v2(M.id3v2_tag(flags=M.ID3_PAD_V2, genre="My Own Genre"))
    assert frames[-1] == ("TCON", "My Own Genre") and pad == 128
    big = M.id3v2_tag(flags=M.ID3_ADD_V2, comment="c" * 70000)
    assert read_v2(big)[0][1] == ("COMM", "c" * 70000)


def test_bad_fields_and_size_query():
    for bad in (dict(year="19x9"), dict(track="one"), dict(genre="148"), dict(genre="-1")):
        with pytest.raises(M.Mp3B200Error):
            M.id3v2_tag(flags=M.ID3_ADD_V2, **bad)
    import ctypes
    import numpy as np
    L = M.lib()
    assert L.mp3b200_id3v1_tag(None, None, 0) == 128
    buf = np.zeros(8, dtype=np.uint8)
    from lamejs_b200.encoder import _id3_struct
    c = _id3_struct({"title": "x"}, M.ID3_ADD_V2, 0, -1, 0)
    need = L.mp3b200_id3v2_tag(ctypes.byref(c), buf.ctypes.data, 8)
    assert need == 10 + (10 + 1 + len(LAME)) + (10 + 1 + 1) and not buf.any()
    L.mp3b200_id3_genre_name.restype = ctypes.c_char_p
    assert L.mp3b200_id3_genre_name(17) == b"Rock"
    assert [L.mp3b200_id3_genre_name(i) is not None for i in (-1, 0, 147, 148)] == [False, True, True, False]
