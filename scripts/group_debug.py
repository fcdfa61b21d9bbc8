import sys
sys.path.insert(0,'/root/repo')
import sleipnir_b200 as sb
N=60
Ts=[4.0+0.25*i for i in range(8)]
r2=sb.multistart("cart_pole",N,Ts)
for T,s in zip(Ts,r2["starts"]):
    Q=sb.Problem("cart_pole",N,T); sq=Q.solve(); print(T,"grouped",s[0],s[2],"alone",sq,len(Q.trace())); Q.close()
