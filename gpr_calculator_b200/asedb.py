"""Minimal reader / writer of the ASE sqlite database rows GP.save / GP.load exchange
(gaussianprocess.py:689-821), usable without ASE.

The reference stores one row per training structure through ``ase.db``: the atoms (numbers, positions,
cell, pbc), ``key_value_pairs`` {dft_energy per atom, dft_fmax} and ``data`` {energy, force[n,3],
energy_in, force_in}.  ASE's sqlite layout (format version 9): raw little-endian blobs for the arrays
(numbers int32, positions / cell float64, pbc as a 3-bit mask) and, for ``data``, a blob made of an int64
offset, the raw arrays and a JSON trailer in which every array is ``{"__ndarray__": [shape, dtype, offset]}``.
Files written here can be opened by ``ase.db.connect`` and files written by ASE can be read here.
"""
import json
import os
import sqlite3
import struct
import time
import uuid

import numpy as np

from .utilities import SimpleAtoms

_SCHEMA = [
    """CREATE TABLE systems (
    id INTEGER PRIMARY KEY AUTOINCREMENT, unique_id TEXT UNIQUE, ctime REAL, mtime REAL, username TEXT,
    numbers BLOB, positions BLOB, cell BLOB, pbc INTEGER, initial_magmoms BLOB, initial_charges BLOB,
    masses BLOB, tags BLOB, momenta BLOB, constraints TEXT, calculator TEXT, calculator_parameters TEXT,
    energy REAL, free_energy REAL, forces BLOB, stress BLOB, dipole BLOB, magmoms BLOB, magmom REAL,
    charges BLOB, key_value_pairs TEXT, data BLOB, natoms INTEGER, fmax REAL, smax REAL, volume REAL,
    mass REAL, charge REAL)""",
    "CREATE TABLE species (Z INTEGER, n INTEGER, id INTEGER, FOREIGN KEY (id) REFERENCES systems(id))",
    "CREATE TABLE keys (key TEXT, id INTEGER, FOREIGN KEY (id) REFERENCES systems(id))",
    "CREATE TABLE text_key_values (key TEXT, value TEXT, id INTEGER, FOREIGN KEY (id) REFERENCES systems(id))",
    "CREATE TABLE number_key_values (key TEXT, value REAL, id INTEGER, FOREIGN KEY (id) REFERENCES systems(id))",
    "CREATE TABLE information (name TEXT, value TEXT)",
    "CREATE INDEX unique_id_index ON systems(unique_id)",
    "CREATE INDEX ctime_index ON systems(ctime)",
    "CREATE INDEX species_index ON species(Z)",
    "CREATE INDEX key_index ON keys(key)",
    "CREATE INDEX number_index ON number_key_values(key)",
]
_YEAR = 31557600.0   # ASE counts ctime in years since 2000-01-01


def _decode_data(blob):
    """int64 offset | raw arrays | JSON trailer  ->  dict with numpy arrays"""
    if blob is None:
        return {}
    if isinstance(blob, str):                      # very old files: plain JSON text
        return json.loads(blob)
    blob = bytes(blob)
    off = struct.unpack("<q", blob[:8])[0]

    def walk(o):
        if isinstance(o, dict):
            if "__ndarray__" in o:
                shape, dtype, start = o["__ndarray__"]
                n = int(np.prod(shape)) if len(shape) else 1
                return np.frombuffer(blob, dtype=np.dtype(dtype), count=n, offset=start).reshape(shape).copy()
            return {k: walk(v) for k, v in o.items()}
        if isinstance(o, list):
            return [walk(v) for v in o]
        return o

    return walk(json.loads(blob[off:].decode()))


def _encode_data(data):
    chunks, pos = [], 8

    def walk(o):
        nonlocal pos
        if isinstance(o, np.ndarray):
            a = np.ascontiguousarray(o)
            if a.dtype == np.int32 or a.dtype == np.bool_:
                a = a.astype(np.int64)
            pad = (-pos) % 8
            chunks.append(b"\0" * pad + a.tobytes())
            start = pos + pad
            pos = start + a.nbytes
            return {"__ndarray__": [list(a.shape), a.dtype.name, start]}
        if isinstance(o, dict):
            return {k: walk(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return [walk(v) for v in o]
        if isinstance(o, (np.integer,)):
            return int(o)
        if isinstance(o, (np.floating,)):
            return float(o)
        if isinstance(o, (np.bool_,)):
            return bool(o)
        return o

    trailer = json.dumps(walk(data), separators=(",", ":")).encode()
    body = b"".join(chunks)
    return struct.pack("<q", 8 + len(body)) + body + trailer


class Row:
    """One database row: atoms + key_value_pairs + data (attribute access like ase.db rows)."""

    def __init__(self, id_, numbers, positions, cell, pbc, key_value_pairs, data):
        self.id = id_
        self.numbers, self.positions, self.cell, self.pbc = numbers, positions, cell, pbc
        self.key_value_pairs = key_value_pairs
        self.data = data

    def toatoms(self):
        return SimpleAtoms(self.numbers, self.positions, self.cell, self.pbc)


def read_rows(filename):
    """Rows of an ASE sqlite database in id order."""
    if not os.path.exists(filename):
        raise FileNotFoundError(filename)
    con = sqlite3.connect("file:%s?mode=ro" % filename, uri=True)
    try:
        cur = con.execute("SELECT id, numbers, positions, cell, pbc, key_value_pairs, data FROM systems ORDER BY id")
        for id_, numbers, positions, cell, pbc, kvp, data in cur:
            numbers = np.frombuffer(numbers, dtype=np.int32).astype(np.int64)
            positions = np.frombuffer(positions, dtype=np.float64).reshape(-1, 3).copy()
            cell = np.frombuffer(cell, dtype=np.float64).reshape(3, 3).copy() if cell is not None else np.zeros((3, 3))
            pbc = np.array([bool(pbc & 1), bool(pbc & 2), bool(pbc & 4)])
            yield Row(id_, numbers, positions, cell, pbc, json.loads(kvp) if kvp else {}, _decode_data(data))
    finally:
        con.close()


def write_rows(filename, rows, append=False):
    """rows: iterable of (atoms, key_value_pairs dict, data dict).  atoms needs numbers / positions / cell / pbc."""
    if not append and os.path.exists(filename):
        os.remove(filename)
    new = not os.path.exists(filename)
    con = sqlite3.connect(filename)
    try:
        if new:
            for stmt in _SCHEMA:
                con.execute(stmt)
            con.execute("INSERT INTO information VALUES ('version', '9')")
        now = (time.time() - 946684800.0) / _YEAR
        for atoms, kvp, data in rows:
            numbers = np.asarray(atoms.numbers, dtype=np.int32)
            positions = np.ascontiguousarray(np.asarray(atoms.positions, dtype=np.float64))
            cell = np.ascontiguousarray(np.asarray(atoms.cell, dtype=np.float64).reshape(3, 3))
            pbc = np.asarray(atoms.pbc, dtype=bool)
            mask = int(pbc[0]) | int(pbc[1]) << 1 | int(pbc[2]) << 2
            kvp = {k: (float(v) if isinstance(v, (float, np.floating)) else v) for k, v in (kvp or {}).items()}
            cur = con.execute(
                "INSERT INTO systems (unique_id, ctime, mtime, username, numbers, positions, cell, pbc, key_value_pairs, "
                "data, natoms, volume, charge) VALUES (?,?,?,?,?,?,?,?,?,?,?,?,?)",
                (uuid.uuid4().hex, now, now, os.environ.get("USER", "gpr"), numbers.tobytes(), positions.tobytes(), cell.tobytes(),
                 mask, json.dumps(kvp), _encode_data(data), len(numbers), abs(float(np.linalg.det(cell))), 0.0))
            rid = cur.lastrowid
            zs, counts = np.unique(numbers, return_counts=True)
            con.executemany("INSERT INTO species VALUES (?,?,?)", [(int(z), int(c), rid) for z, c in zip(zs, counts)])
            con.executemany("INSERT INTO keys VALUES (?,?)", [(k, rid) for k in kvp])
            con.executemany("INSERT INTO number_key_values VALUES (?,?,?)",
                            [(k, float(v), rid) for k, v in kvp.items() if isinstance(v, (int, float))])
        con.commit()
    finally:
        con.close()
