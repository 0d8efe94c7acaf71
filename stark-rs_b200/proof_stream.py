"""Host mirror of the reference's ProofStream (src/stream.rs:4-64) and FiatShamir (src/fiat_shamir.rs:4-25).

Used by the sharded prover (distributed.py), whose per-round orchestration runs on the host; the single-GPU path
assembles the same bytes on the device (csrc/fri.cu).  `serialize` is byte-identical to stream.rs:35-64:
    MerkleRoot      0x00 | 32 bytes
    FieldElement    0x01 | value u64 LE
    FieldElements   0x02 | count u64 LE | values u64 LE ...
    MerklePath      0x03 | count u64 LE | 32-byte hashes ...
"""
import struct


class ProofStream:
    def __init__(self):
        self.objects = []          # (tag, payload)

    def push_merkle_root(self, root):
        assert len(root) == 32
        self.objects.append((0, bytes(root)))

    def push_field_element(self, value):
        self.objects.append((1, int(value)))

    def push_field_elements(self, values):
        self.objects.append((2, [int(v) for v in values]))

    def push_merkle_path(self, hashes):
        self.objects.append((3, [bytes(h) for h in hashes]))

    def push_merkle_path_raw(self, count, raw):
        """a path given as `count` concatenated 32-byte hashes (one bytes object: no per-hash Python objects)"""
        assert len(raw) == 32 * count
        self.objects.append((4, (int(count), bytes(raw))))

    def serialize(self):
        out = bytearray()
        for tag, x in self.objects:
            out.append(tag)
            if tag == 0:
                out += x
            elif tag == 1:
                out += struct.pack("<Q", x)
            elif tag == 2:
                out += struct.pack("<Q", len(x))
                out += struct.pack("<%dQ" % len(x), *x)
            elif tag == 3:
                out += struct.pack("<Q", len(x))
                for h in x:
                    out += h
            else:                      # raw MerklePath: serialised exactly like tag 3
                out[-1] = 3
                out += struct.pack("<Q", x[0])
                out += x[1]
        return bytes(out)


class FiatShamir:
    """Append-only transcript (fiat_shamir.rs:4-25).  `challenge_fn(bytes) -> raw u64` is the product's
    stark_fiat_shamir_challenge (api.fiat_shamir_challenge); the value is NOT reduced mod p (fiat_shamir.rs:21-24)."""

    def __init__(self, challenge_fn, transcript=b""):
        self.transcript = bytearray(transcript)
        self._challenge = challenge_fn

    def absorb(self, data):
        self.transcript += bytes(data)

    def challenge(self):
        return self._challenge(bytes(self.transcript))
