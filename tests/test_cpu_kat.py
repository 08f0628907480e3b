"""6502 known-answer tests from the PUBLIC specification of the NMOS 6502 (MOS programming manual / datasheet
tables) -- not from either emulator: every documented opcode's base cycle count and length, page-cross and
branch timing, binary ADC / SBC results and flags, decimal-mode ADC / SBC on valid BCD operands (result digits and
carry; N / V / Z in decimal mode are implementation lore and are NOT asserted), compares, shifts and rotates,
flag transfers through PHP / PLP.  Both emulator cores must agree with these answers: the oracle console
(oracle/a2600.hpp) and the device core compiled for the host (csrc/emu_core.cuh).  This pins the part of the oracle
that CAN be pinned without ALE: the CPU.  (TIA / RIOT / ALE-layer fidelity stays unpinned, see DESIGN.md.)"""
import ctypes as C

import numpy as np
import pytest

from test_host_emu import libs  # noqa: F401  (fixture: builds + loads both libraries)

# ---- the documented instruction set: opcode -> (mnemonic, mode, bytes, base cycles, +1 on page cross)
# (MOS MCS6500 programming manual, appendix: instruction timing)
DOC = {}


def _add(mn, table):
    for mode, (opc, n, cyc, px) in table.items():
        DOC[opc] = (mn, mode, n, cyc, px)


for mn, base in (("ORA", 0x00), ("AND", 0x20), ("EOR", 0x40), ("ADC", 0x60), ("LDA", 0xA0), ("CMP", 0xC0), ("SBC", 0xE0)):
    _add(mn, {"imm": (base + 0x09, 2, 2, 0), "zp": (base + 0x05, 2, 3, 0), "zpx": (base + 0x15, 2, 4, 0),
              "abs": (base + 0x0D, 3, 4, 0), "abx": (base + 0x1D, 3, 4, 1), "aby": (base + 0x19, 3, 4, 1),
              "izx": (base + 0x01, 2, 6, 0), "izy": (base + 0x11, 2, 5, 1)})
_add("STA", {"zp": (0x85, 2, 3, 0), "zpx": (0x95, 2, 4, 0), "abs": (0x8D, 3, 4, 0), "abx": (0x9D, 3, 5, 0),
             "aby": (0x99, 3, 5, 0), "izx": (0x81, 2, 6, 0), "izy": (0x91, 2, 6, 0)})
for mn, base in (("ASL", 0x00), ("ROL", 0x20), ("LSR", 0x40), ("ROR", 0x60)):
    _add(mn, {"acc": (base + 0x0A, 1, 2, 0), "zp": (base + 0x06, 2, 5, 0), "zpx": (base + 0x16, 2, 6, 0),
              "abs": (base + 0x0E, 3, 6, 0), "abx": (base + 0x1E, 3, 7, 0)})
for mn, base in (("DEC", 0xC0), ("INC", 0xE0)):
    _add(mn, {"zp": (base + 0x06, 2, 5, 0), "zpx": (base + 0x16, 2, 6, 0), "abs": (base + 0x0E, 3, 6, 0), "abx": (base + 0x1E, 3, 7, 0)})
_add("LDX", {"imm": (0xA2, 2, 2, 0), "zp": (0xA6, 2, 3, 0), "zpy": (0xB6, 2, 4, 0), "abs": (0xAE, 3, 4, 0), "aby": (0xBE, 3, 4, 1)})
_add("LDY", {"imm": (0xA0, 2, 2, 0), "zp": (0xA4, 2, 3, 0), "zpx": (0xB4, 2, 4, 0), "abs": (0xAC, 3, 4, 0), "abx": (0xBC, 3, 4, 1)})
_add("STX", {"zp": (0x86, 2, 3, 0), "zpy": (0x96, 2, 4, 0), "abs": (0x8E, 3, 4, 0)})
_add("STY", {"zp": (0x84, 2, 3, 0), "zpx": (0x94, 2, 4, 0), "abs": (0x8C, 3, 4, 0)})
_add("CPX", {"imm": (0xE0, 2, 2, 0), "zp": (0xE4, 2, 3, 0), "abs": (0xEC, 3, 4, 0)})
_add("CPY", {"imm": (0xC0, 2, 2, 0), "zp": (0xC4, 2, 3, 0), "abs": (0xCC, 3, 4, 0)})
_add("BIT", {"zp": (0x24, 2, 3, 0), "abs": (0x2C, 3, 4, 0)})
for mn, opc, cyc in (("CLC", 0x18, 2), ("SEC", 0x38, 2), ("CLI", 0x58, 2), ("SEI", 0x78, 2), ("CLV", 0xB8, 2), ("CLD", 0xD8, 2),
                     ("SED", 0xF8, 2), ("TAX", 0xAA, 2), ("TXA", 0x8A, 2), ("TAY", 0xA8, 2), ("TYA", 0x98, 2), ("TSX", 0xBA, 2),
                     ("TXS", 0x9A, 2), ("INX", 0xE8, 2), ("INY", 0xC8, 2), ("DEX", 0xCA, 2), ("DEY", 0x88, 2), ("NOP", 0xEA, 2),
                     ("PHA", 0x48, 3), ("PHP", 0x08, 3), ("PLA", 0x68, 4), ("PLP", 0x28, 4)):
    DOC[opc] = (mn, "imp", 1, cyc, 0)
for mn, opc in (("BPL", 0x10), ("BMI", 0x30), ("BVC", 0x50), ("BVS", 0x70), ("BCC", 0x90), ("BCS", 0xB0), ("BNE", 0xD0), ("BEQ", 0xF0)):
    DOC[opc] = (mn, "rel", 2, 2, 0)
DOC[0x4C] = ("JMP", "abs", 3, 3, 0); DOC[0x6C] = ("JMP", "ind", 3, 5, 0); DOC[0x20] = ("JSR", "abs", 3, 6, 0)
DOC[0x60] = ("RTS", "imp", 1, 6, 0); DOC[0x40] = ("RTI", "imp", 1, 6, 0); DOC[0x00] = ("BRK", "imp", 1, 7, 0)
assert len(DOC) == 151, len(DOC)          # the documented NMOS 6502 instruction set


class Core(object):
    """One console of either implementation behind the same five calls."""

    def __init__(self, lib, prefix, rom):
        self.L, self.p = lib, prefix
        mk = getattr(lib, prefix + "console_create")
        mk.restype = C.c_void_p
        mk.argtypes = [C.c_char_p, C.c_int]
        self.h = C.c_void_p(mk(rom, len(rom)))
        for name, args in (("console_step", [C.c_void_p, C.c_int]), ("set_ram", [C.c_void_p, C.c_int, C.c_int]),
                           ("get_cpu", [C.c_void_p, C.c_void_p]), ("get_ram", [C.c_void_p, C.c_void_p]), ("destroy", [C.c_void_p])):
            getattr(lib, prefix + name).argtypes = args

    def step(self, n=1):
        getattr(self.L, self.p + "console_step")(self.h, n)

    def set_ram(self, j, v):
        getattr(self.L, self.p + "set_ram")(self.h, j, v)

    def cpu(self):
        out = np.zeros(10, np.int32)
        getattr(self.L, self.p + "get_cpu")(self.h, out.ctypes.data)
        return dict(A=int(out[0]), X=int(out[1]), Y=int(out[2]), SP=int(out[3]), PC=int(out[4]), P=int(out[5]), cycles=int(out[6]))

    def ram(self):
        out = np.zeros(128, np.uint8)
        getattr(self.L, self.p + "get_ram")(self.h, out.ctypes.data)
        return out

    def close(self):
        getattr(self.L, self.p + "destroy")(self.h)


def _rom(code, at=0xF000, fill=0xEA):
    """4K cartridge image with `code` (bytes) at address `at`, reset vector -> $F000."""
    rom = bytearray([fill]) * 4096
    off = at - 0xF000
    rom[off:off + len(code)] = bytes(code)
    rom[0xFFC] = 0x00; rom[0xFFD] = 0xF0
    rom[0xFFE] = 0x00; rom[0xFFF] = 0xF8
    return bytes(rom)


def _cores(libs, rom):
    L, H = libs
    H.he_set_fast_mode(2)        # fast tick checked against the general path on the way
    return [Core(L, "orc_", rom), Core(H, "he_", rom)]


def _run(libs, code, n_instr, ram=None, extra=None):
    """State of both cores after `n_instr` instructions of `code`; asserts they agree with each other first."""
    rom = bytearray(_rom(code))
    for at, data in (extra or {}).items():
        rom[at - 0xF000:at - 0xF000 + len(data)] = bytes(data)
    res = []
    for core in _cores(libs, bytes(rom)):
        for j, v in (ram or {}).items():
            core.set_ram(j & 0x7F, v)
        c0 = core.cpu()
        core.step(n_instr)
        c1 = core.cpu()
        c1["dcycles"] = c1["cycles"] - c0["cycles"]
        c1["ram"] = core.ram()
        res.append(c1)
        core.close()
    assert {k: v for k, v in res[0].items() if k != "ram"} == {k: v for k, v in res[1].items() if k != "ram"}
    assert np.array_equal(res[0]["ram"], res[1]["ram"])
    return res[0]


# operands that make every addressing mode read RIOT RAM ($80-$FF) without crossing a page (X = Y = 1)
def _operand(mode, cross=False):
    if mode in ("imm",):
        return [0x44]
    if mode in ("zp", "zpx", "zpy"):
        return [0x90]
    if mode in ("abs",):
        return [0x90, 0x00]
    if mode in ("abx", "aby"):
        return [0xFF, 0xF0] if cross else [0x90, 0x00]           # $F0FF + 1 crosses into $F100 (cartridge ROM)
    if mode in ("izx",):
        return [0xA0 - 1]                                          # pointer at $A0/$A1 (X = 1)
    if mode == "izy":
        return [0xA0]
    return []


@pytest.mark.parametrize("opc", sorted(o for o, d in DOC.items() if d[1] not in ("rel",) and d[0] not in ("JMP", "JSR", "RTS", "RTI", "BRK")))
def test_documented_length_and_cycles(libs, opc):
    mn, mode, n, cyc, px = DOC[opc]
    # LDX #1 ; LDY #1 ; <instruction>      pointer ($A0) -> $0090, or -> $F0FF for the page-cross case
    for cross in ([False, True] if px else [False]):
        ram = {0xA0: 0xFF if cross else 0x90, 0xA1: 0xF0 if cross else 0x00}
        code = [0xA2, 0x01, 0xA0, 0x01, opc] + _operand(mode, cross)
        before = _run(libs, code, 2, ram)
        after = _run(libs, code, 3, ram)
        assert after["PC"] - before["PC"] == n, (mn, mode)
        assert after["cycles"] - before["cycles"] == cyc + (1 if cross else 0), (mn, mode, cross)


@pytest.mark.parametrize("opc", sorted(o for o, d in DOC.items() if d[1] == "rel"))
def test_branch_timing(libs, opc):
    """not taken 2, taken 3, taken across a page 4; target = PC after the branch + signed offset."""
    mn = DOC[opc][0]
    flag = {"BPL": (0x80, 0), "BMI": (0x80, 1), "BVC": (0x40, 0), "BVS": (0x40, 1), "BCC": (0x01, 0), "BCS": (0x01, 1),
            "BNE": (0x02, 0), "BEQ": (0x02, 1)}[mn]
    for want_taken in (False, True):
        for cross in (False, True):
            bit_set = (flag[1] == 1) == want_taken
            p = 0x20 | (flag[0] if bit_set else 0)
            # LDA #p ; PHA ; PLP ; Bxx   -- the branch sits at $F0F8 so that +$10 crosses into $F1xx
            pre = [0xA9, p, 0x48, 0x28, 0x4C, 0xF8, 0xF0]                 # ... ; JMP $F0F8
            off = 0x10 if cross else 0x02
            st0 = _run(libs, pre, 4, extra={0xF0F8: [opc, off]})
            st1 = _run(libs, pre, 5, extra={0xF0F8: [opc, off]})
            assert st0["PC"] == 0xF0F8
            if want_taken:
                assert st1["PC"] == 0xF0FA + off, (mn, cross)
                assert st1["cycles"] - st0["cycles"] == (4 if cross else 3), (mn, cross)
            else:
                assert st1["PC"] == 0xF0FA and st1["cycles"] - st0["cycles"] == 2, mn


def test_flow_instructions(libs):
    # JSR $F010 (6) ... RTS (6); JMP abs (3); JMP (ind) (5) with the page-wrap bug of the NMOS part
    st = _run(libs, [0x20, 0x10, 0xF0], 1, extra={0xF010: [0x60]})
    assert st["PC"] == 0xF010 and st["dcycles"] == 6 and st["SP"] == 0xFD
    assert st["ram"][0x7F] == 0xF0 and st["ram"][0x7E] == 0x02            # return address - 1, high byte pushed first
    st2 = _run(libs, [0x20, 0x10, 0xF0], 2, extra={0xF010: [0x60]})
    assert st2["PC"] == 0xF003 and st2["dcycles"] == 12 and st2["SP"] == 0xFF
    assert _run(libs, [0x4C, 0x34, 0xF2], 1)["PC"] == 0xF234
    ind = _run(libs, [0x6C, 0xFF, 0xF1], 1, extra={0xF1FF: [0x21], 0xF100: [0xF3], 0xF200: [0xF4]})
    assert ind["PC"] == 0xF321 and ind["dcycles"] == 5                    # high byte from $F100, not $F200
    brk = _run(libs, [0x00], 1)
    assert brk["PC"] == 0xF800 and brk["dcycles"] == 7 and brk["SP"] == 0xFC and (brk["P"] & 0x04)
    assert brk["ram"][0x7F] == 0xF0 and brk["ram"][0x7E] == 0x02 and (brk["ram"][0x7D] & 0x30) == 0x30
    rti = _run(libs, [0x00], 2, extra={0xF800: [0x40]})
    assert rti["PC"] == 0xF002 and rti["dcycles"] == 13 and rti["SP"] == 0xFF


def _flags_program(op_bytes, a, p_in):
    # LDA #p ; PHA ; PLP ; LDA #a ; <op> ; PHP ; PLA (-> A = status) ... A after the op is kept in X first: TAX
    return [0xA9, p_in | 0x20, 0x48, 0x28, 0xA9, a] + op_bytes


def _binary_adc(a, m, c):
    s = a + m + c
    r = s & 0xFF
    v = (~(a ^ m) & (a ^ r) & 0x80) != 0
    return r, int(s > 0xFF), int(v)


def test_adc_sbc_binary(libs):
    rng = np.random.RandomState(6502)
    cases = [(0, 0, 0), (0xFF, 1, 0), (0x7F, 1, 0), (0x80, 0x80, 0), (0x80, 0xFF, 1), (0x7F, 0x7F, 1), (0xFF, 0xFF, 1)]
    cases += [tuple(int(x) for x in rng.randint(0, 256, 2)) + (int(rng.randint(2)),) for _ in range(120)]
    for a, m, c in cases:
        st = _run(libs, _flags_program([0x69, m], a, c), 5)                # ADC #m
        r, co, v = _binary_adc(a, m, c)
        assert st["A"] == r and (st["P"] & 1) == co and ((st["P"] >> 6) & 1) == v, ("ADC", a, m, c)
        assert ((st["P"] >> 7) & 1) == (r >> 7) and ((st["P"] >> 1) & 1) == int(r == 0)
        st = _run(libs, _flags_program([0xE9, m], a, c), 5)                # SBC #m = ADC #~m
        r, co, v = _binary_adc(a, m ^ 0xFF, c)
        assert st["A"] == r and (st["P"] & 1) == co and ((st["P"] >> 6) & 1) == v, ("SBC", a, m, c)


def test_adc_sbc_decimal_on_valid_bcd(libs):
    """Decimal mode, valid BCD operands: the result digits and the carry are what decimal arithmetic says."""
    rng = np.random.RandomState(10)
    bcd = lambda n: ((n // 10) << 4) | (n % 10)
    cases = [(0, 0, 0), (99, 1, 0), (99, 99, 1), (50, 50, 0), (9, 1, 0), (19, 1, 1), (0, 1, 0), (0, 0, 1)]
    cases += [(int(rng.randint(100)), int(rng.randint(100)), int(rng.randint(2))) for _ in range(150)]
    for x, y, c in cases:
        st = _run(libs, _flags_program([0x69, bcd(y)], bcd(x), 0x08 | c), 5)
        s = x + y + c
        assert st["A"] == bcd(s % 100) and (st["P"] & 1) == int(s > 99), ("ADC dec", x, y, c)
        st = _run(libs, _flags_program([0xE9, bcd(y)], bcd(x), 0x08 | c), 5)
        d = x - y - (1 - c)
        assert st["A"] == bcd(d % 100) and (st["P"] & 1) == int(d >= 0), ("SBC dec", x, y, c)


def test_compares_shifts_and_loads(libs):
    rng = np.random.RandomState(3)
    for _ in range(60):
        a, m, c = int(rng.randint(256)), int(rng.randint(256)), int(rng.randint(2))
        st = _run(libs, _flags_program([0xC9, m], a, c), 5)                # CMP #m
        d = (a - m) & 0xFF
        assert st["A"] == a and (st["P"] & 1) == int(a >= m) and ((st["P"] >> 1) & 1) == int(a == m) and (st["P"] >> 7) == (d >> 7)
        for opc, fn in ((0x0A, lambda v, ci: ((v << 1) & 0xFF, v >> 7)), (0x4A, lambda v, ci: (v >> 1, v & 1)),
                        (0x2A, lambda v, ci: (((v << 1) | ci) & 0xFF, v >> 7)), (0x6A, lambda v, ci: ((v >> 1) | (ci << 7), v & 1))):
            st = _run(libs, _flags_program([opc], a, c), 5)
            r, co = fn(a, c)
            assert st["A"] == r and (st["P"] & 1) == co and ((st["P"] >> 1) & 1) == int(r == 0) and (st["P"] >> 7) == (r >> 7), hex(opc)
        st = _run(libs, _flags_program([0x29, m], a, c), 5)                # AND
        assert st["A"] == (a & m) and (st["P"] & 1) == c
        st = _run(libs, _flags_program([0x24, 0x90], a, c), 5, ram={0x90: m})   # BIT zp: N, V from memory, Z from A & M
        assert (st["P"] >> 7) == (m >> 7) and ((st["P"] >> 6) & 1) == ((m >> 6) & 1) and ((st["P"] >> 1) & 1) == int((a & m) == 0)


def test_stack_and_status_transfers(libs):
    # PHP pushes the status with B and bit 5 set; PLP ignores B; PLA sets N / Z
    st = _run(libs, [0xA9, 0xC3, 0x48, 0x28, 0x08, 0x68], 5)
    assert st["A"] == (0xC3 | 0x30) and st["SP"] == 0xFF
    st = _run(libs, [0xA9, 0x00, 0x48, 0xA9, 0x55, 0x68], 4)
    assert st["A"] == 0 and (st["P"] & 0x02) and st["dcycles"] == 2 + 3 + 2 + 4
    # TXS does not touch the flags, TSX does
    st = _run(libs, [0xA2, 0x00, 0x9A, 0xA2, 0x80, 0xBA], 4)
    assert st["X"] == 0 and st["SP"] == 0 and (st["P"] & 0x02)
    # zero-page indexing stays inside page zero
    st = _run(libs, [0xA2, 0x05, 0xB5, 0x90], 2, ram={0x95: 0x5A})
    assert st["A"] == 0x5A
