import importlib, sys
sys.path.insert(0, '.')
mod = importlib.import_module("ntt-based-polynomial-multiplier-fpga_b200")
names = {0:"IMAD",1:"IMAD.HI",2:"IADD",3:"shoup bfly (HI+2IMAD+2IADD)",4:"IMAD.WIDE+LOP",5:"IMAD+IADD pairs(x2)",6:"IMNMX+IADD(x2)",7:"SHFL",8:"IMAD.WIDE.S32 64b addend",9:"plantard-signed bfly (IMAD+WIDE+2IADD)",10:"signed shoup fused bfly",11:"halfword plantard bfly",12:"SHF+IADD(x2)",13:"unsigned plantard bfly (IMAD+HI+2IADD3)"}
for k in sorted(names):
    r = mod.measure_int_peak(k)
    print(f"{k:2d} {names[k]:45s} {r/1e12:8.3f} T/s   per SM-clk @1.965GHz: {r/148/1.965e9:7.2f}")
