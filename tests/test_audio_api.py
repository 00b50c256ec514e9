"""The host-side mirror of needle::audio (needle_b200/audio.py): host-only
behaviour on CPU, the Analyzer -> .needle.dat -> Comparator -> .skip.json flow
on the GPU (reads like needle/src/lib.rs:20-100)."""
import json
import os
import wave

import numpy as np
import pytest

from needle_b200 import audio, synth
from tests import helpers as H


def write_wav(path, pcm, channels=1):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(2)
        w.setframerate(11025)
        w.writeframes(np.asarray(pcm, dtype="<i2").tobytes())


def test_format_time_and_durations():
    assert audio.format_time(0) == "00:00s"
    assert audio.format_time(83_999_999_999) == "01:23s"      # whole seconds truncated (util.rs:8-12)
    assert audio.format_time(3_725_000_000_000) == "62:05s"
    assert audio.duration_from_secs_f32(0.3) == 300_000_012
    assert audio.duration_mul_f32(1_200_000_000_000, 0.5) == 600_000_000_000
    assert audio.duration_mul_f32(1_440_000_000_000, np.float32(1.0) - np.float32(0.25)) == 1_080_000_000_000


def test_with_extension_and_names():
    assert audio._with_extension("/a/b/ep01.mkv", audio.FRAME_HASH_DATA_FILE_NAME) == "/a/b/ep01.needle.dat"
    assert audio._with_extension("/a/b/ep.01.wav", audio.SKIP_FILE_NAME) == "/a/b/ep.01.needle.skip.json"
    assert audio._with_extension("noext", "needle.dat") == "noext.needle.dat"


def test_header_md5_and_discovery(tmp_path):
    rng = np.random.default_rng(0)
    a = tmp_path / "a.wav"
    write_wav(a, rng.integers(-100, 100, 20_000))
    import hashlib
    assert audio.compute_header_md5sum(str(a)) == hashlib.md5(a.read_bytes()[:8192]).hexdigest()
    short = tmp_path / "short.wav"
    write_wav(short, np.zeros(100, np.int16))
    with pytest.raises(audio.NeedleError):
        audio.compute_header_md5sum(str(short))               # read_exact on < 8 KiB
    (tmp_path / "notes.txt").write_text("x")
    (tmp_path / "a.needle.dat").write_bytes(b"\0" * 100)
    sub = tmp_path / "sub"
    sub.mkdir()
    write_wav(sub / "deep.wav", np.zeros(10, np.int16))
    found = audio.find_video_files([str(tmp_path)])
    assert sorted(os.path.basename(p) for p in found) == ["a.wav", "short.wav"]   # one level deep only
    with pytest.raises(audio.PathNotFound):
        audio.find_video_files([str(tmp_path / "nope")])


def test_skip_file_bytes_are_what_serde_json_writes(tmp_path):
    """f32 seconds in their shortest round-tripping form, as serde_json (ryu) prints them
    (comparator.rs:329-354): 12.3, not 12.300000190734863."""
    v = tmp_path / "ep.wav"
    v.write_bytes(b"\0" * 9000)
    audio.Comparator.create_skip_file(str(v), audio.SearchResult(opening=(12_300_000_000, 99_000_000_000)))
    text = (tmp_path / "ep.needle.skip.json").read_text()
    assert text.startswith('{"opening":[12.3,99.0],"ending":null,"md5":"')
    for x, want in ((12.3, "12.3"), (12.0, "12.0"), (0.0, "0.0"), (1e-5, "1e-5"), (3.5e-7, "3.5e-7"), (1e16, "1e16"),
                    (2.6, "2.6"), (1234.5678, "1234.5677")):
        assert audio.json_f32(x) == want
        assert np.float32(float(audio.json_f32(x))) == np.float32(x)


def test_skip_file_roundtrip(tmp_path):
    v = tmp_path / "ep.wav"
    write_wav(v, np.zeros(10_000, np.int16))
    assert not audio.Comparator.check_skip_file(str(v))
    audio.Comparator.create_skip_file(str(v), audio.SearchResult())              # nothing found: no file
    assert not (tmp_path / "ep.needle.skip.json").exists()
    audio.Comparator.create_skip_file(str(v), audio.SearchResult(opening=(10_500_000_000, 99_000_000_000)))
    d = json.loads((tmp_path / "ep.needle.skip.json").read_text())
    assert d["opening"] == [10.5, 99.0] and d["ending"] is None and d["md5"] == audio.compute_header_md5sum(str(v))
    assert audio.Comparator.check_skip_file(str(v))


def test_frame_hashes_file_errors(tmp_path):
    with pytest.raises(audio.FrameHashDataNotFound):
        audio.FrameHashes.from_path(str(tmp_path / "x.needle.dat"))
    bad = tmp_path / "bad.needle.dat"
    bad.write_bytes(b"\x01\0\0\0" + b"\0" * 80)
    with pytest.raises(audio.FrameHashDataInvalidVersion):
        audio.FrameHashes.from_path(str(bad))
    fh = audio.FrameHashes(np.arange(5, dtype=np.uint32), np.arange(5, dtype=np.uint64) * 1000,
                           np.zeros(0, np.uint32), np.zeros(0, np.uint64), 300_000_012, "ab" * 16)
    fh.save(str(tmp_path / "ok.needle.dat"))
    back = audio.FrameHashes.from_path(str(tmp_path / "ok.needle.dat"))
    assert np.array_equal(back.opening_hashes, fh.opening_hashes) and back.md5 == fh.md5
    assert back.hash_duration() == 300_000_012 and back.ending_hashes.size == 0


def test_argument_errors():
    with pytest.raises(audio.AnalyzerMissingPaths):
        audio.Analyzer.from_files([]).run(300_000_012, False, True)
    with pytest.raises(audio.ComparatorMinimumPaths):
        audio.Comparator.from_files(["a"]).run_with_frame_hashes([None])


@pytest.mark.gpu
def test_analyze_then_search_flow(ctx, oracle, tmp_path, capsys):
    """needle analyze --include-endings; needle search --include-endings --write-skip-files."""
    eps = synth.make_pcm_season(4, 5.0, season_seed=11, intro_s=42.0, credits_s=36.0)
    videos = []
    for k, ep in enumerate(eps):
        p = tmp_path / ("ep%02d.wav" % k)
        write_wav(p, np.repeat(ep.pcm, 2) if k % 2 == 0 else np.repeat(ep.pcm, 2), channels=2)   # needle feeds stereo
        videos.append(str(p))
    hd = audio.duration_from_secs_f32(audio.DEFAULT_HASH_DURATION)
    analyzer = audio.Analyzer.from_files(videos, False, False, ctx=ctx).with_include_endings(True)
    frame_hashes = analyzer.run(hd, True, True)
    assert all(os.path.exists(audio._with_extension(v, "needle.dat")) for v in videos)
    # second run reuses the files (md5 of the header matches)
    again = analyzer.run(hd, True, True)
    assert "Skipping analysis" in capsys.readouterr().out
    assert all(np.array_equal(a.opening_hashes, b.opening_hashes) for a, b in zip(frame_hashes, again))
    # hashes agree with the oracle's on the same mono mix
    agree = total = 0
    for ep, fh in zip(eps, frame_hashes):
        a, b, sk = synth.split_segments(ep.pcm)
        wh, wt = oracle.subsample_and_stamp(oracle.fingerprint(a), 2)
        assert np.array_equal(fh.opening_ts_ns, wt)
        agree += int(np.sum(fh.opening_hashes == wh))
        total += wh.size
        eh, et = oracle.subsample_and_stamp(oracle.fingerprint(b), 2, seek_to_ns=sk)
        assert np.array_equal(fh.ending_ts_ns, et)
    assert agree / total >= 0.995
    comparator = audio.Comparator.from_files(videos, ctx=ctx).with_include_endings(True)
    results = comparator.run(False, True, False, True, True)        # from the .needle.dat files
    out = capsys.readouterr().out
    assert len(results) == 4 and "* Opening - " in out and "* Ending - " in out
    for ep, r, v in zip(eps, results, videos):
        assert abs(r.opening[0] / 1e9 - ep.intro_at) < 4.0
        assert abs(r.ending[0] / 1e9 - ep.credits_at) < 4.0
        skip = json.loads(open(audio._with_extension(v, "needle.skip.json")).read())
        # the file holds the shortest decimal of each f32 (serde_json): equal as f32, not as text of the double
        assert [np.float32(x) for x in skip["opening"]] == [np.float32(audio.duration_as_secs_f32(r.opening[0])),
                                                            np.float32(audio.duration_as_secs_f32(r.opening[1]))]
    # bit-exact against the oracle on the GPU's own hashes
    season = H.season_from_lists([f.opening_data() for f in frame_hashes], [f.ending_data() for f in frame_hashes])
    st, want, _ = H.oracle_run(oracle, season, **H.params_kw(include_endings=True))
    assert st == 0
    assert [(r.opening, r.ending) for r in results] == [((w[3], w[4]), (w[5], w[6])) for w in want]
    # skip files now short-circuit the vote (comparator.rs:599-605)
    assert comparator.run(False, False, True, False, True) == []
    # search --analyze analyses without endings -> FrameHashDataNoEnding with include_endings
    with pytest.raises(audio.FrameHashDataNoEnding):
        comparator.run(True, False, False, False, True)
    res2 = audio.Comparator.from_files(videos, ctx=ctx).run(True, False, False, False, True)
    st, want2, _ = H.oracle_run(oracle, season, **H.params_kw(include_endings=False))
    assert [r.opening for r in res2] == [(w[3], w[4]) for w in want2] and all(r.ending is None for r in res2)
