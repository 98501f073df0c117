(* OCaml side of the drop-in (SOURCE ONLY: not compiled in the build image, no OCaml there).

   Keeps the reference's functor signatures: [Curve.S], [Protocol.S],
   [Groth16.Make (C : Curve.S)], [Pinocchio.Make (C)] — only the bodies of the hot functions
   change.  The aliases at the bottom make the README's spelling ([Ecp.Bls12_381],
   [Protocol.Test]) resolve as well as the real one ([Curve.Bls12_381], [Test.Make]).

   The C side is ocaml/zkb200_stubs.c over include/zkb200.h; tests/test_cpu_contract.py compiles the
   stubs against that header.  The Python mirror (zukelang_b200/groth16.py, pinocchio.py) is the
   executable counterpart of this file and is what the parity tests drive. *)

external init : int -> unit = "zkb200_init"
external init_devices : int array -> unit = "zkb200_init_devices"   (* one process, <= 8 GPUs: prove uses them all *)
external table_load : bool -> bytes -> bool -> int64 = "zkb200_table_load"    (* g2?, bases, precompute? *)
external table_msm : bool -> int64 -> bytes -> bytes = "zkb200_table_msm"     (* g2?, handle, 32 n -> 144 | 288 *)
external table_free : int64 -> unit = "zkb200_table_free"
external g1_msm : bytes -> bytes -> bytes = "zkb200_g1_msm"        (* 96 n, 32 n -> 144 *)
external g2_msm : bytes -> bytes -> bytes = "zkb200_g2_msm"        (* 192 n, 32 n -> 288 *)
external g1_sum : bytes -> bytes = "zkb200_g1_sum"
external g2_sum : bytes -> bytes = "zkb200_g2_sum"
external qap_load : bytes -> bytes -> bytes -> bytes -> int -> int -> int64
  = "zkb200_qap_load_bytecode" "zkb200_qap_load_native"
external qap_free : int64 -> unit = "zkb200_qap_free"
external groth16_pk_load : int -> int -> int array -> bytes array -> int * int -> int64
  = "zkb200_groth16_pk_load"
external pinocchio_pk_load : int -> int -> int array -> bytes array -> int * int -> int64
  = "zkb200_pinocchio_pk_load"
external key_free : int64 -> unit = "zkb200_key_free"
external groth16_prove : int64 -> int64 -> bytes -> bytes -> bytes -> bytes = "zkb200_groth16_prove"
external pinocchio_prove : int64 -> int64 -> bytes -> bytes -> bytes = "zkb200_pinocchio_prove"
external pairing_product : bytes -> bytes -> bytes -> bytes = "zkb200_pairing_product"   (* 96n, 192n, n flags|"" -> 576 *)
external gt_mul : bytes -> bytes -> bytes = "zkb200_gt_mul"
external g1_decompress : bytes -> bytes = "zkb200_g1_decompress"   (* 48n -> 96n *)
external g2_decompress : bytes -> bytes = "zkb200_g2_decompress"   (* 96n -> 192n *)

open Zukelang

let cat = Bytes.concat Bytes.empty

(* Device residency (SURVEY.md section 8b: "keys are uploaded once and cached by a device handle stored
   next to pkey").  The OCaml values of the reference (pkey, QAP.t, point lists) are immutable, so a
   device handle is attached to a value by PHYSICAL equality: the first use uploads, later uses of
   the same value find the handle.  A table holds at most [capacity] handles; the least recently
   used one is freed when a new value arrives, and [clear] frees them all (call it before
   zk_shutdown or at exit). *)
module Resident = struct
  type 'a t = { mutable entries : ('a * int64) list; capacity : int; free : int64 -> unit }

  let create ?(capacity = 8) free = { entries = []; capacity; free }

  let rec split n = function
    | x :: xs when n > 0 -> let keep, drop = split (n - 1) xs in (x :: keep, drop)
    | rest -> ([], rest)

  let find_or_load (t : 'a t) (key : 'a) (load : 'a -> int64) : int64 =
    match List.partition (fun (k, _) -> k == key) t.entries with
    | (_, h) :: _, others -> t.entries <- (key, h) :: others; h
    | [], _ ->
        let h = load key in
        let keep, drop = split (t.capacity - 1) t.entries in
        List.iter (fun (_, old) -> t.free old) drop;
        t.entries <- (key, h) :: keep;
        h

  let clear t = List.iter (fun (_, h) -> t.free h) t.entries; t.entries <- []
end

(* [Curve.Bls12_381] with the MSM-shaped members of ExtendMap (curve.ml:79-119) rerouted. *)
module Bls12_381 = struct
  include Curve.Bls12_381

  let fr_bytes (xs : Fr.t list) = cat (List.map Fr.to_bytes xs)

  module G1 = struct
    include Curve.Bls12_381.G1

    let msm (pts : t list) (ks : Fr.t list) : t =
      of_bytes_exn (Bytes.sub (g1_msm (cat (List.map to_bytes pts)) (fr_bytes ks)) 0 96)

    (* Key fields are used again and again (pkey.ti1 for every proof): the point container itself
       — the [Var.Map.t] or the list, by physical equality — owns a resident table.  A table MSM
       takes a prefix of the table, which is exactly apply_powers' "first [length cs] points". *)
    let maps : t Var.Map.t Resident.t = Resident.create table_free
    let lists : t list Resident.t = Resident.create table_free
    let load pts = table_load false (cat (List.map to_bytes pts)) (List.length pts >= 4096)
    let table_msm h (ks : Fr.t list) : t =
      if ks = [] then zero else of_bytes_exn (Bytes.sub (table_msm false h (fr_bytes ks)) 0 96)

    (* curve.ml:91 *)
    let sum_map m f = Var.Map.fold (fun k v acc -> f k v :: acc) m [] |> fun ps ->
      msm ps (List.map (fun _ -> Fr.one) ps)

    (* curve.ml:94-103 *)
    let dot m c =
      if not (Var.Set.equal (Var.Map.domain m) (Var.Map.domain c)) then begin
        prerr_endline "Domain mismatch"; assert false end;
      let bs = Var.Map.bindings m in
      if bs = [] then zero else
      let h = Resident.find_or_load maps m (fun _ -> load (List.map snd bs)) in
      table_msm h (List.map (fun (k, _) -> Var.Infix.(c #! k)) bs)

    (* curve.ml:112-118 *)
    let apply_powers (cs : Fr.t Polynomial.t) xis =
      if List.length cs > List.length xis then invalid_arg "apply_powers";
      if cs = [] then zero else
      table_msm (Resident.find_or_load lists xis load) cs
  end

  module G2 = struct
    include Curve.Bls12_381.G2

    let msm (pts : t list) (ks : Fr.t list) : t =
      of_bytes_exn (Bytes.sub (g2_msm (cat (List.map to_bytes pts)) (fr_bytes ks)) 0 192)

    let maps : t Var.Map.t Resident.t = Resident.create table_free
    let lists : t list Resident.t = Resident.create table_free
    let load pts = table_load true (cat (List.map to_bytes pts)) (List.length pts >= 4096)
    let table_msm h (ks : Fr.t list) : t =
      if ks = [] then zero else of_bytes_exn (Bytes.sub (table_msm true h (fr_bytes ks)) 0 192)

    let dot m c =
      if not (Var.Set.equal (Var.Map.domain m) (Var.Map.domain c)) then begin
        prerr_endline "Domain mismatch"; assert false end;
      let bs = Var.Map.bindings m in
      if bs = [] then zero else
      let h = Resident.find_or_load maps m (fun _ -> load (List.map snd bs)) in
      table_msm h (List.map (fun (k, _) -> Var.Infix.(c #! k)) bs)

    let apply_powers (cs : Fr.t Polynomial.t) xis =
      if List.length cs > List.length xis then invalid_arg "apply_powers";
      if cs = [] then zero else
      table_msm (Resident.find_or_load lists xis load) cs
  end

  (* frees every resident table (before zk_shutdown) *)
  let release_tables () =
    Resident.clear G1.maps; Resident.clear G1.lists; Resident.clear G2.maps; Resident.clear G2.lists
end

(* Verifier side.  A GT value of the library is 576 opaque bytes (NOT blst's GT.t): the device
   verifier keeps its own type and never mixes with [Curve.Bls12_381.GT].  [product] is a sum of
   pairings in the reference's additive notation, [neg] marking the subtracted ones. *)
module Gt_b200 = struct
  type t = bytes
  let zero = let b = Bytes.make 576 '\000' in Bytes.set b 47 '\001'; b
  let ( + ) = gt_mul
  let equal = Bytes.equal
  let product (pairs : (Bls12_381.G1.t * Bls12_381.G2.t * bool) list) : t =
    pairing_product
      (cat (List.map (fun (p, _, _) -> Bls12_381.G1.to_bytes p) pairs))
      (cat (List.map (fun (_, q, _) -> Bls12_381.G2.to_bytes q) pairs))
      (Bytes.of_seq (Seq.map (fun (_, _, n) -> if n then '\001' else '\000') (List.to_seq pairs)))
  let pairing p q = product [ (p, q, false) ]
end

(* Device residency of a QAP.t (QAP.ml:11-16): rows in increasing Var order, coefficients lowest
   degree first, zero padded to n = degree target.  Cached next to the value by the caller. *)
module Device (C : module type of Bls12_381) = struct
  open C
  module QAP = QAP.Make (Fr)
  module Poly = Fr.Poly

  let keys (q : QAP.t) = List.map fst (Var.Map.bindings q.v)

  let pad n p =
    let rec go i = function
      | _ when i = n -> []
      | [] -> Fr.zero :: go (i + 1) []
      | c :: cs -> c :: go (i + 1) cs
    in
    fr_bytes (go 0 p)

  let upload_qap (q : QAP.t) : int64 =
    let n = Poly.degree q.target in
    let flat m = cat (List.map (fun (_, p) -> pad n p) (Var.Map.bindings m)) in
    qap_load (flat q.v) (flat q.w) (flat q.y) (pad (n + 1) q.target) (List.length (keys q)) n

  (* the dense QAP stays on the device for as long as the [QAP.t] is in use (uploaded once) *)
  let qaps : QAP.t Resident.t = Resident.create ~capacity:4 qap_free
  let qap (q : QAP.t) : int64 = Resident.find_or_load qaps q upload_qap

  let index_of ks k =
    let rec go i = function [] -> assert false | x :: xs -> if x = k then i else go (i + 1) xs in
    go 0 ks

  let solution (q : QAP.t) sol = fr_bytes (List.map (fun k -> Var.Infix.(sol #! k)) (keys q))
end

(* Groth16.Make (C).prove, groth16.ml:235-237, with the body moved to the GPU.  keygen / verify are
   the reference's own (include Groth16.Make (C)). *)
module Groth16 (C : module type of Bls12_381) = struct
  include Groth16.Make (C)
  open C
  module D = Device (C)

  let upload_key (pk : pkey) (q : qap) : int64 =
    let ks = D.keys q in
    let n = Fr.Poly.degree q.target in
    let mids = Var.Map.bindings pk.ltd_mid in
    let g1s l = cat (List.map G1.to_bytes l) and g2s l = cat (List.map G2.to_bytes l) in
    groth16_pk_load n (List.length ks)
      (Array.of_list (List.map (fun (k, _) -> D.index_of ks k) mids))
      [| G1.to_bytes pk.a; G1.to_bytes pk.b1; G1.to_bytes pk.d1; G2.to_bytes pk.b2; G2.to_bytes pk.d2;
         g1s (Misclib.List.take n pk.ti1); g2s (Misclib.List.take n pk.ti2); g1s pk.tiztd;
         g1s (List.map snd mids) |]
      (0, 1)

  (* the proving key is uploaded (and its window tables precomputed) on first use and stays
     resident next to [pk]: every later [prove] with the same key is one C call *)
  let keys : pkey Resident.t = Resident.create ~capacity:4 key_free
  let device_key (pk : pkey) (q : qap) : int64 = Resident.find_or_load keys pk (fun pk -> upload_key pk q)
  let release () = Resident.clear keys; Resident.clear D.qaps

  let prove rng (q : qap) (pk : pkey) sol : proof =
    let r = Fr.gen rng in                                  (* groth16.ml:124 — r first *)
    let s = Fr.gen rng in                                  (* groth16.ml:125 *)
    let hq = D.qap q and hk = device_key pk q in
    let out = groth16_prove hk hq (D.solution q sol) (Fr.to_bytes r) (Fr.to_bytes s) in
    { a = G1.of_bytes_exn (Bytes.sub out 0 96);
      b = G2.of_bytes_exn (Bytes.sub out 144 192);
      c = G1.of_bytes_exn (Bytes.sub out 432 96) }

  (* groth16.ml:163-173 with the right-hand pairings moved to the left: one product, one final
     exponentiation.  [ab] must be [Gt_b200.pairing pk.a pk.b2] (the reference's vkey.ab is a blst
     GT.t and cannot be compared with a device GT value). *)
  let verify_b200 ~(ab : Gt_b200.t) w_io (vk : vkey) (pr : proof) =
    Gt_b200.equal ab
      (Gt_b200.product [ (pr.a, pr.b, false); (G1.dot vk.ltgm_io w_io, vk.gm, true); (pr.c, vk.d, true) ])
end

(* Pinocchio.Make (C).{NonZK, ZK}.prove, pinocchio.ml:536-538 / 559-561. *)
module Pinocchio (C : module type of Bls12_381) = struct
  module P = Pinocchio.Make (C)
  open C
  module D = Device (C)

  let upload_key (pk : P.KeyGen.pkey) (q : P.qap) : int64 =
    let ks = D.keys q in
    let n = Fr.Poly.degree q.target in
    let mid = List.map fst (Var.Map.bindings pk.vv) in
    let m1 m = cat (List.map (fun (_, p) -> G1.to_bytes p) (Var.Map.bindings m))
    and m2 m = cat (List.map (fun (_, p) -> G2.to_bytes p) (Var.Map.bindings m)) in
    pinocchio_pk_load n (List.length ks) (Array.of_list (List.map (D.index_of ks) mid))
      [| m1 pk.vv; m1 pk.yy; m1 pk.vav; m1 pk.yay; m1 pk.bvwy; m2 pk.ww; m2 pk.waw;
         cat (List.map G1.to_bytes (Misclib.List.take (n + 1) pk.si)); m1 pk.v_all; m1 pk.w_all;
         G1.to_bytes G1.one; G1.to_bytes pk.vt; G1.to_bytes pk.yt; G1.to_bytes pk.vavt; G1.to_bytes pk.yayt;
         G1.to_bytes pk.vbt; G1.to_bytes pk.wbt; G1.to_bytes pk.ybt; G2.to_bytes pk.wt; G2.to_bytes pk.wawt |]
      (0, 1)

  let unpack out : P.Compute.proof =
    let g1 o = G1.of_bytes_exn (Bytes.sub out o 96) and g2 o = G2.of_bytes_exn (Bytes.sub out o 192) in
    (* vv | ww | yy | h | vavv | waww | yayy | bvwy, 144- and 288-byte point results *)
    { vv = g1 0; ww = g2 144; yy = g1 432; h = g1 576; vavv = g1 720; waww = g2 864; yayy = g1 1152;
      bvwy = g1 1296 }

  let keys : P.KeyGen.pkey Resident.t = Resident.create ~capacity:4 key_free
  let device_key (pk : P.KeyGen.pkey) (q : P.qap) : int64 = Resident.find_or_load keys pk (fun pk -> upload_key pk q)
  let release () = Resident.clear keys; Resident.clear D.qaps

  let prove_with d (q : P.qap) pk sol =
    let hq = D.qap q and hk = device_key pk q in
    unpack (pinocchio_prove hk hq (D.solution q sol) d)

  module NonZK = struct
    include P.NonZK
    let prove _rng q pk sol = prove_with Bytes.empty q pk sol
  end

  module ZK = struct
    include P.ZK
    let prove rng q pk sol =
      let dv = Fr.gen rng in                               (* pinocchio.ml:428 *)
      let dw = Fr.gen rng in                               (* :429 *)
      let dy = Fr.gen rng in                               (* :430 *)
      prove_with (fr_bytes [dv; dw; dy]) q pk sol
  end
end

(* README.md:36-40 spellings.  The code has [Curve.Bls12_381] (curve.mli:56) and [Test.Make] /
   [Test.Make_suites] in library `test` (src/lib/test/test.mli:4-25); [Protocol] holds only the
   signature (protocol.ml:1-21).  These aliases make [Ecp.Bls12_381] and [Protocol.Test (F) (P)]
   resolve too (add `test` to the libraries field of the dune stanza). *)
module Ecp = struct module Bls12_381 = Bls12_381 end
module Protocol = struct
  include Protocol
  module Test_suites = Test.Make_suites     (* before [Test] below shadows the library *)
  module Test = Test.Make
end

(* The reference's own acceptance run (src/groth16/test/main.ml) on the GPU prover:
     module C = Zkb200.Bls12_381
     module G = Zkb200.Groth16 (C)
     let () = Zkb200.init_devices [| 0; 1; 2; 3; 4; 5; 6; 7 |]
     module T = Zkb200.Protocol.Test_suites (C.Fr) (G)                                       *)
