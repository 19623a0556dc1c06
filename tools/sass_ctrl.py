"""Decode the scheduling control fields of sm_100 SASS (cuobjdump -sass): stall count, write / read scoreboard, wait mask.

    python tools/sass_ctrl.py file.cubin [function-substring] [opcode-regex]
Prints one line per instruction whose text matches the regex (default: all): address, stall, yield, write barrier, read barrier,
wait mask, text.  Used to check where ptxas waits for tcgen05.ld (LDTM) results."""
import re, subprocess, sys

def main():
    cubin = sys.argv[1]
    fn = sys.argv[2] if len(sys.argv) > 2 else ""
    rx = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
    out = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout.splitlines()
    infn = False
    i = 0
    while i < len(out):
        l = out[i]
        if "Function :" in l:
            infn = fn in l
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);\s+/\* 0x([0-9a-f]+) \*/", l)
        if infn and m and i + 1 < len(out):
            m2 = re.search(r"/\* 0x([0-9a-f]+) \*/", out[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                stall, yld, wr, rd, wait = (hi >> 41) & 0xf, (hi >> 45) & 1, (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 0x3f
                txt = m.group(2).strip()
                if rx is None or rx.search(txt):
                    print(f"{m.group(1)} st={stall:2d} y={yld} wr={'-' if wr == 7 else wr} rd={'-' if rd == 7 else rd} wait={wait:06b}  {txt}")
                i += 1
        i += 1

if __name__ == "__main__":
    main()
