"""ROS1 bag v2.0 reader / writer of the replay harness (cerberus_b200/rosbag.py): a synthetic robot's sensor streams are
written as the messages the reference subscribes to (sensor_msgs/Imu + sensor_msgs/JointState at 500 Hz with the 16-slot JointState
convention of main.cpp:262-278, the feature tracker's sensor_msgs/PointCloud of main.cpp:200-233), read back without ROS and compared."""
import struct
import numpy as np
from cerberus_b200 import abi, synth, rosbag


def test_bag_round_trip(tmp_path):
    seq = synth.generate_sequence(1, 13, tracked=20, max_len=12, min_len=3)
    path = str(tmp_path / "synthetic_a1.bag")
    msgs = rosbag.sequence_to_messages(seq, 0)
    rosbag.write_bag(path, msgs, chunk_bytes=64 * 1024)
    raw = open(path, "rb").read()
    assert raw.startswith(b"#ROSBAG V2.0\n")
    # the bag header record is padded to 4096 bytes and names the index position, connection and chunk counts
    (hl,) = struct.unpack_from("<I", raw, 13)
    hdr = rosbag._parse_fields(raw[17:17 + hl])
    assert hdr["op"] == b"\x03" and struct.unpack("<I", hdr["conn_count"])[0] == 3 and struct.unpack("<I", hdr["chunk_count"])[0] > 3
    (dl,) = struct.unpack_from("<I", raw, 17 + hl)
    assert 13 + 4 + hl + 4 + dl == 13 + 4096
    idx = struct.unpack("<Q", hdr["index_pos"])[0]
    assert 4096 < idx < len(raw)
    # every message comes back, in time order, with its topic and type
    got = list(rosbag.read_bag(path))
    assert len(got) == len(msgs) and all(a[0] == b[0] and abs(a[1] - b[2]) < 1e-8 for a, b in zip(got, msgs))
    imu = [m for t, _, m in got if t == rosbag.IMU_TOPIC]; leg = [m for t, _, m in got if t == rosbag.LEG_TOPIC]
    assert len(imu) == len(leg) == 12 * 33 + 1
    assert len(leg[0]["position"]) == 16 and len(leg[0]["velocity"]) == 16 and len(leg[0]["effort"]) == 16
    # the replay inputs reconstructed from the bag equal the sequence they were written from
    frames, smp = rosbag.load_sequence(path)
    assert len(frames) == seq.n_frames and len(smp) == 12 * 33 + 1
    flat = seq.samples[0].reshape(-1)
    for name in ("acc", "gyr", "phi", "dphi", "c"):
        assert np.abs(smp[name][1:] - flat[name]).max() == 0.0, name
    assert np.abs(smp["dt"][1:] - flat["dt"]).max() < 1e-8 and np.abs(smp["acc"][0] - seq.first[0, 0]["acc"]).max() == 0.0
    for i, fr in enumerate(frames):
        img = seq.images[i][0]
        assert (fr["ids"] == img["ids"]).all() and (fr["has1"] == img["has1"]).all()
        assert np.abs(fr["pts0"] - img["pts0"].astype(np.float32)).max() < 1e-12                  # PointCloud carries float32
        assert np.abs(fr["pts1"][fr["has1"]] - img["pts1"][img["has1"]].astype(np.float32)).max() < 1e-12
    # topic filter and bz2 chunks
    only = list(rosbag.read_bag(path, topics={rosbag.FEATURE_TOPIC}))
    assert len(only) == seq.n_frames and only[0][2]["points"].shape[1] == 3


def test_bz2_chunk_and_raw_mode(tmp_path):
    import bz2
    path = str(tmp_path / "one.bag")
    body = rosbag.serialize_imu(7, 12.5, [1.0, 2.0, 3.0], [0.1, 0.2, 0.3])
    rosbag.write_bag(path, [("/imu", "sensor_msgs/Imu", 12.5, body)])
    raw = bytearray(open(path, "rb").read())
    # rewrite the single chunk as a bz2 chunk (what `rosbag compress` produces)
    pos = 13 + 4096
    (hl,) = struct.unpack_from("<I", raw, pos); h = rosbag._parse_fields(bytes(raw[pos + 4:pos + 4 + hl]))
    (dl,) = struct.unpack_from("<I", raw, pos + 4 + hl); data = bytes(raw[pos + 8 + hl:pos + 8 + hl + dl])
    comp = bz2.compress(data)
    rec = rosbag._record([("op", bytes([5])), ("compression", b"bz2"), ("size", struct.pack("<I", len(data)))], comp)
    out = bytes(raw[:pos]) + rec
    p2 = str(tmp_path / "bz2.bag"); open(p2, "wb").write(out)
    (topic, t, m), = list(rosbag.read_bag(p2))
    assert topic == "/imu" and abs(t - 12.5) < 1e-9 and m["header"]["seq"] == 7 and (m["linear_acceleration"] == [1.0, 2.0, 3.0]).all() and (m["angular_velocity"] == [0.1, 0.2, 0.3]).all()
    (topic, t, (typ, rawmsg)), = list(rosbag.read_bag(p2, decode=False))
    assert typ == "sensor_msgs/Imu" and rawmsg == body


def test_replay_from_a_bag_matches_replay_from_the_sequence(tmp_path):
    """bag -> load_sequence -> sequence_from_bag -> ReplayDriver (oracle arm) == the same replay fed from the synthetic sequence directly (with
    the float32 rounding of the feature tracker's PointCloud message applied to it too: the reference receives its features as float32)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cerberus_b200 import estimator
    from oracle_lib import OracleOps
    seq = synth.generate_sequence(1, 14, tracked=16, max_len=12, min_len=3)
    path = str(tmp_path / "replay.bag")
    rosbag.write_bag(path, rosbag.sequence_to_messages(seq, 0))
    frames, smp = rosbag.load_sequence(path)
    bseq = rosbag.sequence_from_bag(frames, smp, dict(p_g=seq.p_g[0], R_g=seq.R_g[0], v_g=seq.v_g[0], tic_g=seq.tic_g[0], ric_g=seq.ric_g[0]))
    assert bseq.samples.shape == seq.samples.shape
    cfg = abi.default_config(); cfg.max_batch, cfg.max_features, cfg.max_obs, cfg.max_num_iterations = 1, 128, 128 * 11, 6
    pcfg = abi.default_preint_config()
    for img in seq.images:                     # the same float32 rounding on the direct arm: a 1e-8 input difference grows ~20 x per chained frame
        img[0]["pts0"] = img[0]["pts0"].astype(np.float32).astype(np.float64); img[0]["pts1"] = img[0]["pts1"].astype(np.float32).astype(np.float64)
    seq.samples["dt"] = bseq.samples["dt"]      # stamps are stored with nanosecond resolution: dt = stamp differences, like main.cpp:284-300 forms it
    a = estimator.ReplayDriver(OracleOps(cfg), cfg, pcfg, 1, max_features=64).run(seq)
    b = estimator.ReplayDriver(OracleOps(cfg), cfg, pcfg, 1, max_features=64).run(bseq)
    Pa, Ra = a.poses(); Pb, Rb = b.poses()
    assert Pa.shape == Pb.shape and Pa.shape[1] == 4
    assert np.abs(Pa - Pb).max() < 1e-9 and np.abs(Ra - Rb).max() < 1e-9, (np.abs(Pa - Pb).max(), np.abs(Ra - Rb).max())
    csv = str(tmp_path / "vilo.csv")
    estimator.write_csv(csv, b.est[0], pcfg)
    assert len(open(csv).read().strip().split("\n")) == 4
