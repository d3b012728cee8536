import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/oracle')
import numpy as np, pathlib, tempfile
import cpelnano_b200 as cpn
import test_wgbs_bam as T
lib=cpn.Library()
for kw in (dict(), dict(depth=20,p_high=1.0), dict(depth=4, p_high=0.7)):
    d=pathlib.Path(tempfile.mkdtemp())
    fasta,bam=T.synthetic_wgbs(d,**kw)
    cfg=cpn.CpelNanoConfig()
    chrom=cpn.genome.load_genome(fasta,lib)['chrS']
    part=chrom.chr_part_native(lib,cfg.size_est_reg,cfg.min_grp_dist,None)
    regs=[r for r in chrom.region_structs(lib,part,1) if len(r.cpg_pos)>9]
    calls=T.native_calls(cpn,lib,bam,'chrS',[(r.chrst,r.chrend) for r in regs],[np.asarray(r.cpg_pos,dtype=np.int64) for r in regs])
    cpg_off=np.concatenate([[0],np.cumsum([len(r.cpg_pos) for r in regs])]).astype(np.int64)
    read_off=np.concatenate([[0],np.cumsum([c.shape[0] for c in calls])]).astype(np.int64)
    flat=np.concatenate([c.reshape(-1) for c in calls])
    rho=np.concatenate([r.rho_n for r in regs]); dn=np.concatenate([r.dn for r in regs])
    eng=cpn.default_engine()
    for r in regs: eng.get_nls_reg_info(r,cfg)
    sub_off=np.concatenate([[0],np.cumsum([r.nls_rgs.num for r in regs])]).astype(np.int64)
    sp=np.concatenate([r.nls_rgs.cpg_p for r in regs]); sq=np.concatenate([r.nls_rgs.cpg_q for r in regs])
    phi0=cpn.simulations.gen_phi0s(np.random.default_rng(11),len(regs),10)
    for iters in (20,100):
        out=eng.lib.analyze_wgbs_batch(cpg_off,rho,dn,read_off,flat,phi0,10,sub_off,sp,sq,iters)
        print(kw, 'iters',iters,'M',np.diff(read_off),'pick',out['pick'],'Q',out['Q'],'counters',out['counters'].tolist(),'phi',out['phi_hat'].round(3).tolist())
