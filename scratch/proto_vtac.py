"""numpy prototype of the iterative, phi-factored VTAC algorithm used by the CUDA kernel."""
import sys, numpy as np
sys.path.insert(0,'/root/repo')
from oracle import oracle as O
from scipy import special as sp

def a_plus(n,m):
    return 0.0 if abs(m)>n or n<0 else -np.sqrt((n+m+1)*(n-m+1)/((2*n+1)*(2*n+3)))
def a_minus(n,m):
    return 0.0 if abs(m)>n or n<0 else np.sqrt((n+m)*(n-m)/((2*n+1)*(2*n-1)))
def b_plus(n,m):
    return 0.0 if abs(m)>n or n<0 else np.sqrt((n+m+2)*(n+m+1)/((2*n+1)*(2*n+3)))
def b_minus(n,m):
    return 0.0 if abs(m)>n or n<0 else np.sqrt((n-m)*(n-m-1)/((2*n+1)*(2*n-1)))

def proto(R,k,NM,regular):
    r,the,phi=R
    L=2*NM
    z=k*r
    n=np.arange(L+1)
    zl = sp.spherical_jn(n,z) if regular else sp.spherical_jn(n,z)+1j*sp.spherical_yn(n,z)
    x=np.cos(the)
    # level storage: dict[(m)] -> array[(lam,kap)]
    def idx(l,kk): return l*(l+1)+kk
    lev={}  # lev[n][m] = array
    # seeds
    E=np.zeros((L+1)**2,complex)
    for l in range(L+1):
        for kk in range(-l,l+1):
            N=sp.sph_harm_y(l,abs(kk),the,0.0).real
            sgn=(-1)**l if kk>=0 else (-1)**(l+kk)
            E[idx(l,kk)]=np.sqrt(4*np.pi)*sgn*N*zl[l]
    G={ (0,0):E }
    def get(arr,Lmax,l,kk):
        if l<0 or abs(kk)>l or l>Lmax: return 0.0
        return arr[idx(l,kk)]
    for nn in range(1,NM+1):
        Ln=L-nn
        for m in range(0,nn+1):
            out=np.zeros((Ln+1)**2,complex)
            for l in range(Ln+1):
                for kk in range(-l,l+1):
                    if m==nn:
                        src=G[(nn-1,nn-1)]
                        v=(get(src,Ln+1,l-1,kk-1)*b_plus(l-1,kk-1)+get(src,Ln+1,l+1,kk-1)*b_minus(l+1,kk-1))/b_plus(nn-1,nn-1)
                    else:
                        s1=G[(nn-1,m)]
                        v=get(s1,Ln+1,l-1,kk)*a_plus(l-1,kk)+get(s1,Ln+1,l+1,kk)*a_minus(l+1,kk)
                        if nn-2>=m:
                            v-=get(G[(nn-2,m)],Ln+2,l,kk)*a_minus(nn-1,m)
                        v/=a_plus(nn-1,m)
                    out[idx(l,kk)]=v
            G[(nn,m)]=out
    def beta(nn,mu,l,kk):
        # phase-free: returns R-part incl. sign; full beta = exp(i(mu-kk)phi)*this
        if l<0 or abs(kk)>l or abs(mu)>nn: return 0.0
        if mu>=0: return G[(nn,mu)][idx(l,kk)]
        return (-1)**(mu+kk)*G[(nn,-mu)][idx(l,-kk)]
    N=NM*(NM+2)
    A=np.zeros((N,N),complex); B=np.zeros((N,N),complex)
    for nn in range(1,NM+1):
        for m in range(-nn,nn+1):
            p=nn*(nn+1)-m-1
            for l in range(1,NM+1):
                for kk in range(-l,l+1):
                    q=l*(l+1)-kk-1
                    ph=np.exp(1j*(m-kk)*phi)
                    f=0.5/np.sqrt(l*(l+1)*nn*(nn+1))
                    c0=2*kk*m; c1=np.sqrt((nn-m)*(nn+m+1)*(l-kk)*(l+kk+1)); c2=np.sqrt((nn+m)*(nn-m+1)*(l+kk)*(l-kk+1))
                    A[p,q]=ph*f*(c0*beta(nn,m,l,kk)+c1*beta(nn,m+1,l,kk+1)+c2*beta(nn,m-1,l,kk-1))
                    fb=-0.5j*np.sqrt((2*l+1)/((2*l-1)*l*(l+1)*nn*(nn+1)))
                    d0=2*m*np.sqrt((l-kk)*(l+kk)); d1=np.sqrt(max(0,(nn-m)*(nn+m+1)*(l-kk)*(l-kk-1))); d2=np.sqrt(max(0,(nn+m)*(nn-m+1)*(l+kk)*(l+kk-1)))
                    B[p,q]=ph*fb*(d0*beta(nn,m,l-1,kk)+d1*beta(nn,m+1,l-1,kk+1)-d2*beta(nn,m-1,l-1,kk-1))
    return A,B

if __name__=="__main__":
    k=2*np.pi/800e-9
    for R in ([190e-9,0.9,2.2],[120e-9,2.4,-1.0],[700e-9,np.pi/2,0.0],[300e-9,0.0,0.0],[300e-9,np.pi,0.0]):
        for reg in (False,True):
            NM=5
            A,B=proto(R,k,NM,reg)
            Ao,Bo=O.coupling(R,k,NM, (not reg))   # Coupling ctor flag is inverted
            print(R,reg,"relerr A %.2e B %.2e"%(np.linalg.norm(A-Ao)/np.linalg.norm(Ao),np.linalg.norm(B-Bo)/np.linalg.norm(Bo)))
