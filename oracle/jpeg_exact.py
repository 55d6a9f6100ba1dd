"""TEST INFRASTRUCTURE ONLY -- integer restatement of the baseline-JPEG round trip Pillow performs for
`img.save(buf, format="JPEG", quality=q, subsampling=...)` + `Image.open(buf)` (jpeg_compress, svd.ipynb#c1:L20-44 and
0409_method.ipynb#c0:L44-62), i.e. of libjpeg-turbo (third-party, not under /root/reference; 3.1.x behind Pillow 12 here)
without its lossless entropy coder.  Restated from the library's published algorithm, file by file:
  jcparam.c   jpeg_quality_scaling / jpeg_add_quant_table(force_baseline)      -> qtables
  jccolor.c   rgb_ycc_convert (16-bit fixed point)                              -> rgb2ycc
  jcsample.c  h2v2_downsample (bias 1,2,1,2,...)                               -> down_h2v2
  jfdctint.c  jpeg_fdct_islow (LL&M, CONST_BITS 13, PASS1_BITS 2)               -> fdct_1d
  jcdctmgr.c  quantize: round-half-away division by 8*Q                         -> plane_roundtrip
  jidctint.c  jpeg_idct_islow                                                   -> idct_1d
  jdsample.c  h2v2_fancy_upsample (triangle filter, replicated edges)           -> up_h2v2_fancy
  jdcolor.c   ycc_rgb_convert                                                   -> ycc2rgb
Pinned: bit-exact against Pillow's own round trip for every quality and both subsampling modes on noise, smooth and
hard-edged images of arbitrary size (tests/test_oracle.py::test_jpeg_exact_restatement_vs_pillow), MCU edge padding
included.  The GPU kernels (csrc/jpeg_exact.cu) cover sizes without padding and are checked against this file and Pillow.
"""
import io, numpy as np
from PIL import Image

QY = np.array([16,11,10,16,24,40,51,61,12,12,14,19,26,58,60,55,14,13,16,24,40,57,69,56,14,17,22,29,51,87,80,62,
               18,22,37,56,68,109,103,77,24,35,55,64,81,104,113,92,49,64,78,87,103,121,120,101,72,92,95,98,112,100,103,99],dtype=np.int64).reshape(8,8)
QC = np.array([17,18,24,47,99,99,99,99,18,21,26,66,99,99,99,99,24,26,56,99,99,99,99,99,47,66,99,99,99,99,99,99]+[99]*32,dtype=np.int64).reshape(8,8)

def qtables(quality):
    q = max(1, min(100, int(quality)))
    scale = 5000 // q if q < 50 else 200 - 2 * q
    f = lambda t: np.clip((t * scale + 50) // 100, 1, 255)
    return f(QY), f(QC)

F = dict(f0298=2446, f0390=3196, f0541=4433, f0765=6270, f0899=7373, f1175=9633, f1501=12299, f1847=15137, f1961=16069, f2053=16819, f2562=20995, f3072=25172)
CONST_BITS, PASS1_BITS = 13, 2
def descale(x, n): return (x + (1 << (n - 1))) >> n

def fdct_1d(d, first):
    # d: [..., 8] int64
    d0,d1,d2,d3,d4,d5,d6,d7 = [d[..., i] for i in range(8)]
    tmp0=d0+d7; tmp7=d0-d7; tmp1=d1+d6; tmp6=d1-d6; tmp2=d2+d5; tmp5=d2-d5; tmp3=d3+d4; tmp4=d3-d4
    tmp10=tmp0+tmp3; tmp13=tmp0-tmp3; tmp11=tmp1+tmp2; tmp12=tmp1-tmp2
    out=[None]*8
    if first:
        out[0]=(tmp10+tmp11)<<PASS1_BITS; out[4]=(tmp10-tmp11)<<PASS1_BITS; sh=CONST_BITS-PASS1_BITS
    else:
        out[0]=descale(tmp10+tmp11,PASS1_BITS); out[4]=descale(tmp10-tmp11,PASS1_BITS); sh=CONST_BITS+PASS1_BITS
    z1=(tmp12+tmp13)*F['f0541']
    out[2]=descale(z1+tmp13*F['f0765'],sh); out[6]=descale(z1+tmp12*(-F['f1847']),sh)
    z1=tmp4+tmp7; z2=tmp5+tmp6; z3=tmp4+tmp6; z4=tmp5+tmp7; z5=(z3+z4)*F['f1175']
    tmp4=tmp4*F['f0298']; tmp5=tmp5*F['f2053']; tmp6=tmp6*F['f3072']; tmp7=tmp7*F['f1501']
    z1=z1*(-F['f0899']); z2=z2*(-F['f2562']); z3=z3*(-F['f1961']); z4=z4*(-F['f0390'])
    z3=z3+z5; z4=z4+z5
    out[7]=descale(tmp4+z1+z3,sh); out[5]=descale(tmp5+z2+z4,sh); out[3]=descale(tmp6+z2+z3,sh); out[1]=descale(tmp7+z1+z4,sh)
    return np.stack(out,-1)

def idct_1d(c, first):
    i0,i1,i2,i3,i4,i5,i6,i7=[c[..., i] for i in range(8)]
    z2=i2; z3=i6
    z1=(z2+z3)*F['f0541']; tmp2=z1+z3*(-F['f1847']); tmp3=z1+z2*F['f0765']
    z2=i0; z3=i4
    tmp0=(z2+z3)<<CONST_BITS; tmp1=(z2-z3)<<CONST_BITS
    tmp10=tmp0+tmp3; tmp13=tmp0-tmp3; tmp11=tmp1+tmp2; tmp12=tmp1-tmp2
    tmp0=i7; tmp1=i5; tmp2=i3; tmp3=i1
    z1=tmp0+tmp3; z2=tmp1+tmp2; z3=tmp0+tmp2; z4=tmp1+tmp3; z5=(z3+z4)*F['f1175']
    tmp0=tmp0*F['f0298']; tmp1=tmp1*F['f2053']; tmp2=tmp2*F['f3072']; tmp3=tmp3*F['f1501']
    z1=z1*(-F['f0899']); z2=z2*(-F['f2562']); z3=z3*(-F['f1961']); z4=z4*(-F['f0390'])
    z3=z3+z5; z4=z4+z5
    tmp0=tmp0+z1+z3; tmp1=tmp1+z2+z4; tmp2=tmp2+z2+z3; tmp3=tmp3+z1+z4
    sh = CONST_BITS-PASS1_BITS if first else CONST_BITS+PASS1_BITS+3
    o=[descale(tmp10+tmp3,sh),descale(tmp11+tmp2,sh),descale(tmp12+tmp1,sh),descale(tmp13+tmp0,sh),
       descale(tmp13-tmp0,sh),descale(tmp12-tmp1,sh),descale(tmp11-tmp2,sh),descale(tmp10-tmp3,sh)]
    return np.stack(o,-1)

def plane_roundtrip(p, q):
    """p: [H,W] uint8-valued int64 plane (H,W multiples of 8); q: [8,8] quant table -> decoded plane"""
    H,W=p.shape
    b=(p-128).reshape(H//8,8,W//8,8).transpose(0,2,1,3)          # [by,bx,r,c]
    t=fdct_1d(b,True)                                            # rows: along c
    t=fdct_1d(t.transpose(0,1,3,2),False).transpose(0,1,3,2)     # columns: along r
    div=(q<<3)
    a=np.abs(t); coef=np.sign(t)*((a+(div>>1))//div)
    c=coef*q
    w=idct_1d(c.transpose(0,1,3,2),True).transpose(0,1,3,2)      # pass 1: columns
    o=idct_1d(w,False)                                            # pass 2: rows
    o=np.clip(o+128,0,255)
    return o.transpose(0,2,1,3).reshape(H,W)

def fix(x): return int(x*65536+0.5)
def rgb2ycc(rgb):
    r,g,b=[rgb[...,i].astype(np.int64) for i in range(3)]
    half=1<<15; off=128<<16
    y=(fix(0.29900)*r+fix(0.58700)*g+fix(0.11400)*b+half)>>16
    cb=(-fix(0.16874)*r-fix(0.33126)*g+fix(0.50000)*b+off+half-1)>>16
    cr=(fix(0.50000)*r-fix(0.41869)*g-fix(0.08131)*b+off+half-1)>>16
    return y,cb,cr
def ycc2rgb(y,cb,cr):
    half=1<<15
    x=cr-128; r=y+((fix(1.40200)*x+half)>>16)
    xb=cb-128; b=y+((fix(1.77200)*xb+half)>>16)
    g=y+((-fix(0.34414)*xb+half + -fix(0.71414)*x)>>16)
    return np.stack([np.clip(r,0,255),np.clip(g,0,255),np.clip(b,0,255)],-1)
def down_h2v2(p):
    H,W=p.shape
    s=p[0::2,0::2]+p[0::2,1::2]+p[1::2,0::2]+p[1::2,1::2]
    bias=np.tile(np.array([1,2],dtype=np.int64),W//4+1)[:W//2]
    return (s+bias[None,:])>>2
def up_h2v2_fancy(p):
    h,w=p.shape
    up=np.vstack([p[:1],p[:-1]]); dn=np.vstack([p[1:],p[-1:]])
    out=np.zeros((2*h,2*w),dtype=np.int64)
    for v,far in ((0,up),(1,dn)):
        cs=3*p+far                                            # thiscolsum per column
        last=np.hstack([cs[:,:1],cs[:,:-1]]); nxt=np.hstack([cs[:,1:],cs[:,-1:]])
        e=(3*cs+last+8)>>4; o=(3*cs+nxt+7)>>4
        e[:,0]=(cs[:,0]*4+8)>>4; o[:,-1]=(cs[:,-1]*4+7)>>4
        out[v::2,0::2]=e; out[v::2,1::2]=o
    return out

def roundtrip_rgb(img, quality, sub420):
    """img: [H, W, 3] uint8, any size.  Sizes that are not a multiple of the MCU (16 with 4:2:0, 8 with 4:4:4) are padded the
    way libjpeg does it: the encoder replicates the last row / column of the colour-converted planes up to the MCU boundary
    (jcprepct.c expand_bottom_edge, jcsample.c expand_right_edge); the decoder upsamples and colour-converts only the real
    ceil(H/2) x ceil(W/2) chroma samples (edge cases of the triangle filter at the REAL edge) and crops to H x W."""
    H, W = img.shape[:2]
    mcu = 16 if sub420 else 8
    Hp, Wp = -(-H // mcu) * mcu, -(-W // mcu) * mcu
    qy, qc = qtables(quality)
    y, cb, cr = rgb2ycc(img)
    # columns: replicated at full resolution up to the MCU boundary BEFORE downsampling (jcsample.c expand_right_edge);
    # rows: full resolution only up to a multiple of the vertical sampling factor (jcprepct.c), then every component's
    # (downsampled) plane is padded to the iMCU height by replicating ITS last row (expand_bottom_edge on the output)
    He = H + (H & 1) if sub420 else H
    padw = lambda p: np.pad(p, ((0, He - H), (0, Wp - W)), mode="edge")
    y, cb, cr = padw(y), padw(cb), padw(cr)
    if sub420:
        cb, cr = down_h2v2(cb), down_h2v2(cr)
    padh = lambda p, rows: np.pad(p, ((0, rows - p.shape[0]), (0, 0)), mode="edge")
    y = padh(y, Hp)
    cb, cr = padh(cb, Hp // 2 if sub420 else Hp), padh(cr, Hp // 2 if sub420 else Hp)
    y = plane_roundtrip(y, qy); cb = plane_roundtrip(cb, qc); cr = plane_roundtrip(cr, qc)
    if sub420:
        ch, cw = -(-H // 2), -(-W // 2)
        # jdsample.c jinit_upsampler: the triangle filter is only used when the downsampled width exceeds 2, else replication
        up = up_h2v2_fancy if cw > 2 else (lambda p: np.repeat(np.repeat(p, 2, axis=0), 2, axis=1))
        cb, cr = up(cb[:ch, :cw]), up(cr[:ch, :cw])
    return ycc2rgb(y[:H, :W], cb[:H, :W], cr[:H, :W]).astype(np.uint8)

def pil_roundtrip(img, quality, mode, sub=None):
    buf=io.BytesIO(); kw={}
    if sub is not None: kw['subsampling']=sub
    Image.fromarray(img,mode=mode).save(buf,format='JPEG',quality=quality,**kw); buf.seek(0)
    return np.array(Image.open(buf).convert(mode))

