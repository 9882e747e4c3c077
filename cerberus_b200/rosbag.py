"""ROS1 bag (format version 2.0) reader for the replay harness -- no ROS needed.

What the reference reads from a bag (README.md:91-133, src/main.cpp:200-278, 454-470):
  IMU_TOPIC   sensor_msgs/Imu          500 Hz   linear_acceleration, angular_velocity
  LEG_TOPIC   sensor_msgs/JointState   500 Hz   16 slots: position[0..11] / velocity[0..11] joint angles / rates, velocity[12..15] planned contact
                                                flags, effort[12..15] foot force sensor readings (main.cpp:262-278); approximate-time synchronised with the IMU
  /feature_tracker/feature  sensor_msgs/PointCloud   points (x, y, z = 1) + channels id, camera id, u, v, vx, vy (main.cpp:200-233)
  IMAGE0/1_TOPIC  sensor_msgs/Image    15 Hz    consumed by the feature tracker (frontend: out of scope here; images are returned undecoded)

Format (http://wiki.ros.org/Bags/Format/2.0): "#ROSBAG V2.0\\n", then records  <header_len u32><header><data_len u32><data>,  header = fields
<len u32><name>=<value>;  op 0x03 bag header, 0x05 chunk (compression none / bz2; data = records), 0x07 connection, 0x02 message data,
0x04 index data, 0x06 chunk info.  The reader scans the records sequentially (indices are not needed) and yields (topic, stamp, message).
A small writer (uncompressed chunks) produces files for the tests and for exporting synthetic sequences to ROS tools.
"""
import bz2
import struct
import numpy as np
from . import abi

MAGIC = b"#ROSBAG V2.0\n"
OP_MSG, OP_BAG_HEADER, OP_INDEX, OP_CHUNK, OP_CHUNK_INFO, OP_CONNECTION = 2, 3, 4, 5, 6, 7
MD5 = {"sensor_msgs/Imu": "6a62c6daae103f4ff57a132d6f95cec2", "sensor_msgs/JointState": "3066dcd76a6cfaef579bd0f34173e9fd",
       "sensor_msgs/PointCloud": "d8e9c3f5afbdd8a130fd1d2763945fca", "sensor_msgs/Image": "060021388200f6f0f447d0fcd9c64743"}


# ------------------------------------------------------------------------------------------------ record level
def _parse_fields(buf):
    out, p = {}, 0
    while p < len(buf):
        (n,) = struct.unpack_from("<I", buf, p); p += 4
        k, _, v = buf[p:p + n].partition(b"=")
        out[k.decode()] = v; p += n
    return out


def _records(buf, pos=0, end=None):
    end = len(buf) if end is None else end
    while pos + 8 <= end:
        (hl,) = struct.unpack_from("<I", buf, pos); pos += 4
        hdr = _parse_fields(buf[pos:pos + hl]); pos += hl
        (dl,) = struct.unpack_from("<I", buf, pos); pos += 4
        yield hdr, buf[pos:pos + dl]
        pos += dl


def read_bag(path, topics=None, decode=True):
    """Yield (topic, t [s], msg) in file order (= record time order inside each chunk).  msg: a dict for the message types above when
    decode=True, else (type, raw bytes)."""
    data = open(path, "rb").read()
    if not data.startswith(MAGIC):
        raise ValueError(f"{path}: not a ROS bag v2.0")
    conns = {}

    def handle(hdr, body):
        op = hdr["op"][0]
        if op == OP_CONNECTION:
            cid = struct.unpack("<I", hdr["conn"])[0]
            ch = _parse_fields(body)
            conns[cid] = (hdr["topic"].decode(), ch.get("type", b"").decode())
        elif op == OP_MSG:
            cid = struct.unpack("<I", hdr["conn"])[0]
            topic, typ = conns[cid]
            if topics is not None and topic not in topics:
                return None
            secs, nsecs = struct.unpack("<II", hdr["time"])
            return topic, secs + 1e-9 * nsecs, (deserialize(typ, body) if decode else (typ, bytes(body)))
        return None

    for hdr, body in _records(data, len(MAGIC)):
        op = hdr["op"][0]
        if op == OP_CHUNK:
            comp = hdr["compression"].decode()
            if comp == "bz2":
                body = bz2.decompress(body)
            elif comp != "none":
                raise NotImplementedError(f"chunk compression '{comp}' (the reference's bags are uncompressed, README.md:99)")
            for h2, b2 in _records(body):
                r = handle(h2, b2)
                if r is not None:
                    yield r
        else:
            r = handle(hdr, body)
            if r is not None:
                yield r


# ------------------------------------------------------------------------------------------------ ROS1 serialisation of the message types
class _Rd:
    def __init__(self, b): self.b, self.p = b, 0
    def u32(self): (v,) = struct.unpack_from("<I", self.b, self.p); self.p += 4; return v
    def string(self): n = self.u32(); s = bytes(self.b[self.p:self.p + n]).decode(errors="replace"); self.p += n; return s
    def arr(self, dt, n=None):
        n = self.u32() if n is None else n
        a = np.frombuffer(self.b, dtype=dt, count=n, offset=self.p).copy(); self.p += a.nbytes; return a
    def header(self):
        seq, secs, nsecs = struct.unpack_from("<III", self.b, self.p); self.p += 12
        return {"seq": seq, "stamp": secs + 1e-9 * nsecs, "frame_id": self.string()}


def deserialize(typ, body):
    r = _Rd(body)
    if typ == "sensor_msgs/Imu":
        h = r.header()
        q = r.arr("<f8", 4); r.arr("<f8", 9); w = r.arr("<f8", 3); r.arr("<f8", 9); a = r.arr("<f8", 3)
        return {"header": h, "orientation": q, "angular_velocity": w, "linear_acceleration": a}
    if typ == "sensor_msgs/JointState":
        h = r.header()
        names = [r.string() for _ in range(r.u32())]
        return {"header": h, "name": names, "position": r.arr("<f8"), "velocity": r.arr("<f8"), "effort": r.arr("<f8")}
    if typ == "sensor_msgs/PointCloud":
        h = r.header()
        pts = r.arr("<f4", 3 * r.u32()).reshape(-1, 3)
        ch = {}
        for _ in range(r.u32()):
            name = r.string(); ch[name] = r.arr("<f4")
        return {"header": h, "points": pts, "channels": ch}
    if typ == "sensor_msgs/Image":
        h = r.header()
        height, width = r.u32(), r.u32(); enc = r.string(); big = r.b[r.p]; r.p += 1; step = r.u32()
        return {"header": h, "height": height, "width": width, "encoding": enc, "is_bigendian": big, "step": step, "data": r.arr("u1")}
    return {"type": typ, "raw": bytes(body)}


def _ser_header(seq, stamp, frame_id=""):
    secs = int(np.floor(stamp)); nsecs = int(round((stamp - secs) * 1e9))
    if nsecs >= 1000000000: secs, nsecs = secs + 1, nsecs - 1000000000
    f = frame_id.encode()
    return struct.pack("<III", seq, secs, nsecs) + struct.pack("<I", len(f)) + f


def serialize_imu(seq, stamp, acc, gyr):
    z9 = np.zeros(9).tobytes()
    return _ser_header(seq, stamp) + np.array([0, 0, 0, 1.0]).tobytes() + z9 + np.asarray(gyr, "<f8").tobytes() + z9 + np.asarray(acc, "<f8").tobytes() + z9


def serialize_joint_state(seq, stamp, position, velocity, effort, names=None):
    names = names or [f"j{k}" for k in range(len(position))]
    out = _ser_header(seq, stamp) + struct.pack("<I", len(names))
    for n in names:
        out += struct.pack("<I", len(n)) + n.encode()
    for a in (position, velocity, effort):
        a = np.asarray(a, "<f8"); out += struct.pack("<I", a.size) + a.tobytes()
    return out


def serialize_feature_cloud(seq, stamp, points, channels):
    """The feature tracker's message: points [k, 3] + channels (dict name -> [k] values), main.cpp:200-233 reads channels 0..5 = id, camera, u, v, vx, vy."""
    pts = np.asarray(points, "<f4")
    out = _ser_header(seq, stamp, "world") + struct.pack("<I", pts.shape[0]) + pts.tobytes() + struct.pack("<I", len(channels))
    for name, v in channels.items():
        v = np.asarray(v, "<f4"); out += struct.pack("<I", len(name)) + name.encode() + struct.pack("<I", v.size) + v.tobytes()
    return out


# ------------------------------------------------------------------------------------------------ writer (uncompressed chunks)
def _field(name, value): b = name.encode() + b"=" + value; return struct.pack("<I", len(b)) + b
def _record(fields, data): h = b"".join(_field(k, v) for k, v in fields); return struct.pack("<I", len(h)) + h + struct.pack("<I", len(data)) + data


def write_bag(path, messages, chunk_bytes=768 * 1024):
    """messages: iterable of (topic, type, t [s], serialized bytes), time ordered."""
    conns, chunks, cur, cur_info = {}, [], [], None

    def conn_record(cid, topic, typ):
        ch = _field("topic", topic.encode()) + _field("type", typ.encode()) + _field("md5sum", MD5.get(typ, "*").encode()) + _field("message_definition", b"")
        return _record([("op", bytes([OP_CONNECTION])), ("conn", struct.pack("<I", cid)), ("topic", topic.encode())], ch)

    def tstamp(t):
        secs = int(np.floor(t)); nsecs = int(round((t - secs) * 1e9))
        if nsecs >= 1000000000: secs, nsecs = secs + 1, nsecs - 1000000000
        return struct.pack("<II", secs, nsecs)

    def flush():
        nonlocal cur, cur_info
        if cur:
            chunks.append((b"".join(cur), cur_info)); cur, cur_info = [], None

    for topic, typ, t, body in messages:
        if cur_info is None:
            cur_info = {"start": t, "end": t, "counts": {}}
        if (topic, typ) not in conns:
            conns[(topic, typ)] = len(conns)
            cur.append(conn_record(conns[(topic, typ)], topic, typ))
        cid = conns[(topic, typ)]
        cur.append(_record([("op", bytes([OP_MSG])), ("conn", struct.pack("<I", cid)), ("time", tstamp(t))], body))
        cur_info["end"] = t; cur_info["counts"][cid] = cur_info["counts"].get(cid, 0) + 1
        if sum(len(c) for c in cur) >= chunk_bytes:
            flush()
    flush()
    with open(path, "wb") as f:
        f.write(MAGIC)
        hdr_pos = f.tell()
        f.write(b"\0" * 4096)
        infos = []
        for data, info in chunks:
            pos = f.tell()
            f.write(_record([("op", bytes([OP_CHUNK])), ("compression", b"none"), ("size", struct.pack("<I", len(data)))], data))
            infos.append((pos, info))
        index_pos = f.tell()
        for (topic, typ), cid in conns.items():
            f.write(conn_record(cid, topic, typ))
        for pos, info in infos:
            body = b"".join(struct.pack("<II", c, k) for c, k in info["counts"].items())
            f.write(_record([("op", bytes([OP_CHUNK_INFO])), ("ver", struct.pack("<I", 1)), ("chunk_pos", struct.pack("<Q", pos)), ("start_time", tstamp(info["start"])),
                             ("end_time", tstamp(info["end"])), ("count", struct.pack("<I", len(info["counts"])))], body))
        f.seek(hdr_pos)
        h = b"".join(_field(k, v) for k, v in [("op", bytes([OP_BAG_HEADER])), ("index_pos", struct.pack("<Q", index_pos)), ("conn_count", struct.pack("<I", len(conns))),
                                               ("chunk_count", struct.pack("<I", len(chunks)))])
        pad = 4096 - 4 - len(h) - 4
        f.write(struct.pack("<I", len(h)) + h + struct.pack("<I", pad) + b" " * pad)


# ------------------------------------------------------------------------------------------------ bag <-> replay harness
IMU_TOPIC, LEG_TOPIC, FEATURE_TOPIC = "/hardware_a1/imu", "/hardware_a1/joint_foot", "/feature_tracker/feature"


def sequence_to_messages(seq, w=0, t0=100.0):
    """One robot of a synthetic SynthSequence as the messages the reference subscribes to (IMU + JointState at 500 Hz, the feature tracker's
    PointCloud at the camera rate), time ordered."""
    S = seq.samples.shape[2]
    frame_dt = S * seq.dt
    out, k = [], 0
    def leg_msg(smp, t, k):
        pos = np.zeros(16); vel = np.zeros(16); eff = np.zeros(16)
        pos[:12] = smp["phi"]; vel[:12] = smp["dphi"]; vel[12:] = smp["c"]; eff[12:] = smp["c"] * 100.0
        return (LEG_TOPIC, "sensor_msgs/JointState", t, serialize_joint_state(k, t, pos, vel, eff))
    for i in range(seq.n_frames):
        t = t0 + i * frame_dt
        if i == 0:
            f = seq.first[w, 0]
            out.append((IMU_TOPIC, "sensor_msgs/Imu", t, serialize_imu(k, t, f["acc"], f["gyr"]))); out.append(leg_msg(f, t, k)); k += 1
        img = seq.images[i][w]
        n = len(img["ids"]); both = np.nonzero(img["has1"])[0]
        ids = np.concatenate([img["ids"], img["ids"][both]]).astype(np.float32); cam = np.concatenate([np.zeros(n), np.ones(both.size)])
        p = np.concatenate([img["pts0"], img["pts1"][both]])
        out.append((FEATURE_TOPIC, "sensor_msgs/PointCloud", t, serialize_feature_cloud(i, t, p[:, 0:3], {"id": ids, "cam": cam, "u": p[:, 3], "v": p[:, 4], "vx": p[:, 5], "vy": p[:, 6]})))
        if i + 1 < seq.n_frames:
            for j in range(S):
                smp = seq.samples[w, i, j]; ts = t + (j + 1) * seq.dt
                out.append((IMU_TOPIC, "sensor_msgs/Imu", ts, serialize_imu(k, ts, smp["acc"], smp["gyr"]))); out.append(leg_msg(smp, ts, k)); k += 1
    out.sort(key=lambda m: m[2])
    return out


def load_sequence(path, imu_topic=IMU_TOPIC, leg_topic=LEG_TOPIC, feature_topic=FEATURE_TOPIC, sync_slop=0.001):
    """Read one robot's streams back: (frames, samples) with
       frames : list of dict(t, ids, pts0 [k,7], has1, pts1 [k,7]) in the layout ReplayDriver.step takes (main.cpp:200-233: x, y, z, u, v, vx, vy)
       samples: structured array (t, abi.sample_dtype fields) of the IMU + leg pairs, approximate-time synchronised like main.cpp:454-470
                (dt = difference of consecutive stamps; slots per main.cpp:262-278)."""
    imu, leg, frames = [], [], []
    for topic, t, m in read_bag(path, topics={imu_topic, leg_topic, feature_topic}):
        if topic == imu_topic: imu.append((m["header"]["stamp"], m["linear_acceleration"], m["angular_velocity"]))
        elif topic == leg_topic: leg.append((m["header"]["stamp"], m["position"], m["velocity"], m["effort"]))
        else:
            ch = list(m["channels"].values())
            ids, cam = ch[0].astype(np.int64), ch[1].astype(np.int64)
            p7 = np.concatenate([m["points"].astype(np.float64), np.stack([c.astype(np.float64) for c in ch[2:6]], axis=1)], axis=1)
            left = cam == 0
            lid = ids[left]; order = {int(v): k for k, v in enumerate(lid)}
            pts1 = np.zeros((lid.size, 7)); has1 = np.zeros(lid.size, dtype=bool)
            for q in np.nonzero(~left)[0]:
                k = order.get(int(ids[q]))
                if k is not None: pts1[k] = p7[q]; has1[k] = True
            frames.append({"t": m["header"]["stamp"], "ids": lid, "pts0": p7[left], "has1": has1, "pts1": pts1})
    dt = np.dtype([("t", "<f8")] + [(n, abi.sample_dtype.fields[n][0]) for n in abi.sample_dtype.names])
    out = np.zeros(len(imu), dtype=dt)
    j, n = 0, 0
    for (t, acc, gyr) in imu:
        while j + 1 < len(leg) and abs(leg[j + 1][0] - t) <= abs(leg[j][0] - t): j += 1
        if not leg or abs(leg[j][0] - t) > sync_slop: continue
        _, pos, vel, eff = leg[j]
        o = out[n]
        o["t"] = t; o["acc"] = acc; o["gyr"] = gyr; o["phi"] = pos[:12]; o["dphi"] = vel[:12]; o["c"] = vel[12:16]
        o["dt"] = t - out[n - 1]["t"] if n else 0.0
        n += 1
    return frames, out[:n]


class BagSequence:
    """What ReplayDriver.seed / run read of a sequence (see synth.SynthSequence), rebuilt from a bag's streams for ONE robot."""
    pass


def sequence_from_bag(frames, samples, init, time_slop=1e-4):
    """frames / samples: the outputs of load_sequence; init: dict(p_g [F,3], R_g [F,3,3], v_g [F,3], tic_g [2,3], ric_g [2,3,3]) -- the states the
    window is seeded with (the reference gets them from its initialisation, which is out of scope here).  Samples are cut into inter-frame
    intervals by the camera stamps: interval k = samples with stamp in (t_k, t_k+1]; first[k] = the sample at t_k (processIMULeg's acc_0 / gyr_0 ...)."""
    nF = len(frames)
    seq = BagSequence()
    seq.n, seq.n_frames = 1, nF
    ft = np.array([f["t"] for f in frames]); st = samples["t"]
    names = abi.sample_dtype.names
    def strip(a):
        out = np.zeros(len(a), dtype=abi.sample_dtype)
        for n in names: out[n] = a[n]
        return out
    first = np.zeros((1, nF), dtype=abi.sample_dtype); per = []
    for k in range(nF):
        i0 = int(np.argmin(np.abs(st - ft[k])))
        if abs(st[i0] - ft[k]) > time_slop: raise ValueError(f"no IMU / leg sample at camera stamp {ft[k]:.6f}")
        first[0, k] = strip(samples[i0:i0 + 1])[0]
        if k + 1 < nF:
            sel = (st > ft[k] + time_slop) & (st <= ft[k + 1] + time_slop)
            per.append(strip(samples[sel]))
    S = max(len(p) for p in per)
    if any(len(p) != S for p in per): seq.samples = [per]                      # ragged intervals: list indexing [w][k]
    else: seq.samples = np.stack(per)[None]
    seq.first = first
    seq.images = [[{"ids": f["ids"], "pts0": f["pts0"], "has1": f["has1"], "pts1": f["pts1"]}] for f in frames]
    seq.p_g, seq.R_g, seq.v_g = init["p_g"][None], init["R_g"][None], init["v_g"][None]
    seq.tic_g, seq.ric_g = init["tic_g"][None], init["ric_g"][None]
    return seq
