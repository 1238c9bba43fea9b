"""Injectable stand-in for corr.katcp_wrapper.FpgaClient (the reference's hardware transport,
ROACH_Setup.py:112-115): records every register / BRAM access; never talks to hardware."""
import struct


class FakeRoach:
    def __init__(self):
        self.log = []            # (op, name, value)
        self.ints = {}
        self.mem = {}

    def is_connected(self):
        return True

    def write_int(self, name, value, *a, **k):
        self.log.append(('write_int', name, int(value)))
        self.ints[name] = int(value)

    def read_int(self, name, *a, **k):
        self.log.append(('read_int', name, None))
        return self.ints.get(name, 0)

    def write(self, name, data, offset=0):
        data = bytes(data)
        self.log.append(('write', name, len(data)))
        self.mem[name] = data

    def read(self, name, size, offset=0):
        self.log.append(('read', name, size))
        d = self.mem.get(name, b'')
        d = d[offset:offset + size]
        return d + b'\x00' * (size - len(d))

    def writes(self, name):
        return [v for op, n, v in self.log if n == name and op == 'write_int']
