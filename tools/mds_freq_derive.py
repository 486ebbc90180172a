import itertools, random
C=[17,15,41,16,2,28,13,13,39,18,34,20]
def mds(x):
    y=[sum(C[i]*x[(i+r)%12] for i in range(12)) for r in range(12)]
    y[0]+=8*x[0]
    return y
# Good-Thomas: j = 3a+4b mod 12
idx=lambda a,b:(3*a+4*b)%12
# DFT over a of C: Chat[k][b] = sum_a C[idx(a,b)] * i^(a k)
I=1j
Chat=[[sum(C[idx(a,b)]*(I**(a*k)) for a in range(4)) for b in range(3)] for k in range(4)]
for k in range(4): print(k,[Chat[k][b]/4 for b in range(3)])
M=0xFFFFFFFF
def freq_mds(x):
    # x: 12 ints (any), arithmetic exact here (python ints); returns circulant part only
    X0=[0]*3;X2=[0]*3;X1r=[0]*3;X1i=[0]*3
    for b in range(3):
        x0,x1,x2,x3=[x[idx(a,b)] for a in range(4)]
        p=x0+x2;q=x1+x3;r=x0-x2;s=x1-x3
        X0[b]=p+q;X2[b]=p-q;X1r[b]=r;X1i[b]=s      # X1 = r + i s   (sum x_a i^a = x0 + i x1 - x2 - i x3)
    # Yhat[k][b] = sum_b' Chat[-k][b'] * Xhat[k][b+b']
    K0=[int((Chat[0][b]/4).real) for b in range(3)]
    K2=[int((Chat[2][b]/4).real) for b in range(3)]
    K1=[Chat[3][b]*2/4 for b in range(3)]   # Chat[-1] = Chat[3]; doubled
    Y0=[sum(K0[bp]*X0[(b+bp)%3] for bp in range(3)) for b in range(3)]
    Y2=[sum(K2[bp]*X2[(b+bp)%3] for bp in range(3)) for b in range(3)]
    Y1=[sum(K1[bp]*complex(X1r[(b+bp)%3],X1i[(b+bp)%3]) for bp in range(3)) for b in range(3)]
    y=[0]*12
    for b in range(3):
        # y[a] = (1/4) sum_k Yhat[k] i^(-a k) ; with Y0,Y2 already /4 and Y1 = 2*Yhat1/4: y = Y0 + (-1)^a Y2 + Re(Y1 * i^(-a))
        for a in range(4):
            v=Y0[b]+((-1)**a)*Y2[b]+(Y1[b]*(I**(-a))).real
            y[idx(a,b)]=int(round(v))
    return y,K0,K2,K1
x=[random.randrange(1<<22) for _ in range(12)]
y,K0,K2,K1=freq_mds(x)
ref=mds(x); ref[0]-=8*x[0]
print(y==ref,K0,K2,K1)
