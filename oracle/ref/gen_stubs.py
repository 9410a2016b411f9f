"""TEST INFRASTRUCTURE ONLY.  Emit trapping stubs for the symbols the partially linked reference objects leave
undefined (code that exists in the reference but is unreachable on the DiffDFSPH/Akinci2012 path and therefore not
compiled).  Calling one prints its name and aborts, so nothing can silently run through a stub."""
import subprocess, sys

so, out = sys.argv[1], sys.argv[2]
syms = []
for line in subprocess.run(["nm", "-u", so], capture_output=True, text=True, check=True).stdout.splitlines():
    parts = line.split()
    if len(parts) != 2 or parts[0] not in ("U",):
        continue
    s = parts[1]
    if "@" in s:  # versioned libc / libstdc++ / libgomp symbols
        continue
    if s.startswith(("_ZN3SPH", "_ZNK3SPH", "_ZN9Utilities", "_ZNK9Utilities", "_ZTIN3SPH", "_ZTVN3SPH", "_ZN3MD5", "te_", "_ZN10Discregrid")):
        syms.append(s)
with open(out, "w") as f:
    f.write("\t.text\n")
    f.write("\t.section .rodata\nstub_msg:\n\t.string \"oracle/_ref: call into a reference symbol that is not built on this path: %s\\n\"\n\t.text\n")
    for i, s in enumerate(syms):
        f.write(f"\t.section .rodata\nstub_name_{i}:\n\t.string \"{s}\"\n\t.text\n")
        f.write(f"\t.globl {s}\n\t.type {s}, @function\n{s}:\n")
        f.write(f"\tleaq stub_msg(%rip), %rdi\n\tleaq stub_name_{i}(%rip), %rsi\n\txorl %eax, %eax\n\tcall printf@PLT\n\tcall abort@PLT\n")
    f.write('\t.section .note.GNU-stack,"",@progbits\n')
print(f"{len(syms)} stubs")
