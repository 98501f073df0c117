(* OCaml side of the drop-in (SOURCE ONLY: not compiled in the build image, no OCaml there).

   Keeps the reference's functor signatures: [Curve.S], [Protocol.S],
   [Groth16.Make (C : Curve.S)], [Pinocchio.Make (C)] — only the bodies of the hot functions
   change.  The aliases at the bottom make the README's spelling ([Ecp.Bls12_381],
   [Protocol.Test]) resolve as well as the real one ([Curve.Bls12_381], [Test.Make]).

   The reference seals its protocol modules: [Groth16.Make (C) : Protocol.S] (groth16.mli:3-6) and
   [Pinocchio.Make (C).{NonZK, ZK} : Protocol.S] (pinocchio.mli:3-15) keep pkey / vkey / proof
   ABSTRACT and export only their yojson converters.  The functors below therefore read a key
   through [yojson_of_pkey] (records as objects, points as the raw compressed bytes in a JSON string:
   curve.ml:199,208), decompress the points on the device ([g1_decompress], one call per field) and
   build proofs with [proof_of_yojson] from the compressed halves of the results — once per key: the
   device handle stays resident next to the key.  Witness scalars and (r, s) take the fast way:
   [Curve.Bls12_381 : S with type Fr.t = Bls12_381.Fr.t ...] (curve.mli:56-60) keeps the carrier
   types equal to the opam library's, so [L.Fr.to_bytes] applies to them directly.

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

module L = Bls12_381                  (* the opam library itself: byte encodings of Fr / G1 / G2 values *)
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

let fr_bytes (xs : L.Fr.t list) = cat (List.map L.Fr.to_bytes xs)

(* [Curve.Bls12_381] with the MSM-shaped members of ExtendMap (curve.ml:79-119) rerouted.  The
   signature [Curve.G] does not export to_bytes / of_bytes_exn; the opam library's own functions
   apply because the carrier types are equal (curve.mli:56-60). *)
module Bls12_381 = struct
  include Curve.Bls12_381

  module G1 = struct
    include Curve.Bls12_381.G1

    let msm (pts : t list) (ks : Fr.t list) : t =
      L.G1.of_bytes_exn (Bytes.sub (g1_msm (cat (List.map L.G1.to_bytes pts)) (fr_bytes ks)) 0 96)

    (* Key fields are used again and again (pkey.ti1 for every proof): the point container itself
       — the [Var.Map.t] or the list, by physical equality — owns a resident table.  A table MSM
       takes a prefix of the table, which is exactly apply_powers' "first [length cs] points". *)
    let maps : t Var.Map.t Resident.t = Resident.create table_free
    let lists : t list Resident.t = Resident.create table_free
    let load pts = table_load false (cat (List.map L.G1.to_bytes pts)) (List.length pts >= 4096)
    let resident_msm h (ks : Fr.t list) : t =
      if ks = [] then zero else L.G1.of_bytes_exn (Bytes.sub (table_msm false h (fr_bytes ks)) 0 96)

    (* curve.ml:91 *)
    let sum_map m f = Var.Map.fold (fun k v acc -> f k v :: acc) m [] |> fun ps ->
      if ps = [] then zero else msm ps (List.map (fun _ -> Fr.one) ps)

    (* curve.ml:94-103 *)
    let dot m c =
      if not (Var.Set.equal (Var.Map.domain m) (Var.Map.domain c)) then begin
        prerr_endline "Domain mismatch"; assert false end;
      let bs = Var.Map.bindings m in
      if bs = [] then zero else
      let h = Resident.find_or_load maps m (fun _ -> load (List.map snd bs)) in
      resident_msm h (List.map (fun (k, _) -> Var.Infix.(c #! k)) bs)

    (* curve.ml:112-118 *)
    let apply_powers (cs : Fr.t Polynomial.t) xis =
      if List.length cs > List.length xis then invalid_arg "apply_powers";
      if cs = [] then zero else
      resident_msm (Resident.find_or_load lists xis load) cs
  end

  module G2 = struct
    include Curve.Bls12_381.G2

    let msm (pts : t list) (ks : Fr.t list) : t =
      L.G2.of_bytes_exn (Bytes.sub (g2_msm (cat (List.map L.G2.to_bytes pts)) (fr_bytes ks)) 0 192)

    let maps : t Var.Map.t Resident.t = Resident.create table_free
    let lists : t list Resident.t = Resident.create table_free
    let load pts = table_load true (cat (List.map L.G2.to_bytes pts)) (List.length pts >= 4096)
    let resident_msm h (ks : Fr.t list) : t =
      if ks = [] then zero else L.G2.of_bytes_exn (Bytes.sub (table_msm true h (fr_bytes ks)) 0 192)

    let dot m c =
      if not (Var.Set.equal (Var.Map.domain m) (Var.Map.domain c)) then begin
        prerr_endline "Domain mismatch"; assert false end;
      let bs = Var.Map.bindings m in
      if bs = [] then zero else
      let h = Resident.find_or_load maps m (fun _ -> load (List.map snd bs)) in
      resident_msm h (List.map (fun (k, _) -> Var.Infix.(c #! k)) bs)

    let apply_powers (cs : Fr.t Polynomial.t) xis =
      if List.length cs > List.length xis then invalid_arg "apply_powers";
      if cs = [] then zero else
      resident_msm (Resident.find_or_load lists xis load) cs
  end

  (* frees every resident table (before zk_shutdown) *)
  let release_tables () =
    Resident.clear G1.maps; Resident.clear G1.lists; Resident.clear G2.maps; Resident.clear G2.lists
end

(* Verifier side.  A GT value of the library is 576 opaque bytes (NOT blst's GT.t): the device
   verifier keeps its own type and never mixes with [Curve.Bls12_381.GT].  [product] is a sum of
   pairings in the reference's additive notation, [neg] marking the subtracted ones; points are
   uncompressed bytes (96 / 192). *)
module Gt_b200 = struct
  type t = bytes
  let zero = let b = Bytes.make 576 '\000' in Bytes.set b 47 '\001'; b
  let ( + ) = gt_mul
  let equal = Bytes.equal
  let product_bytes (pairs : (bytes * bytes * bool) list) : t =
    pairing_product
      (cat (List.map (fun (p, _, _) -> p) pairs))
      (cat (List.map (fun (_, q, _) -> q) pairs))
      (Bytes.of_seq (Seq.map (fun (_, _, n) -> if n then '\001' else '\000') (List.to_seq pairs)))
  let product (pairs : (L.G1.t * L.G2.t * bool) list) : t =
    product_bytes (List.map (fun (p, q, n) -> (L.G1.to_bytes p, L.G2.to_bytes q, n)) pairs)
  let pairing p q = product [ (p, q, false) ]
end

(* The shapes ppx_yojson_conv gives the reference's key and proof records (groth16.ml:24-43,110-114;
   pinocchio.ml:37-75,195-208): records are objects, lists are arrays, a ['a Var.Map.t] is the array of
   its bindings [[var, a], ...] in increasing key order (var.ml:38-40,66-68), and a point is a JSON
   string holding its compressed bytes (curve.ml:199,208). *)
module Json = struct
  type t = Yojson.Safe.t
  let field (j : t) name : t =
    match j with
    | `Assoc l -> (try List.assoc name l with Not_found -> failwith ("zkb200: key field " ^ name ^ " is missing"))
    | _ -> failwith "zkb200: a record was expected"
  let bytes_of : t -> bytes = function `String s -> Bytes.of_string s | _ -> failwith "zkb200: a byte string was expected"
  let list_of : t -> t list = function `List l -> l | _ -> failwith "zkb200: a list was expected"
  let point j name = bytes_of (field j name)                               (* compressed *)
  let points j name = cat (List.map bytes_of (list_of (field j name)))     (* 'a list, compressed, concatenated *)
  let bindings j name : (Var.t * bytes) list =                             (* 'a Var.Map.t *)
    List.map (function `List [ v; p ] -> (Var.t_of_yojson v, bytes_of p) | _ -> failwith "zkb200: a binding was expected")
      (list_of (field j name))
  let of_point (b : bytes) : t = `String (Bytes.to_string b)
end

(* compressed -> uncompressed on the device, one call per field (curve.ml:201,210 over a batch) *)
let g1_raw (comp : bytes) = if Bytes.length comp = 0 then Bytes.empty else g1_decompress comp
let g2_raw (comp : bytes) = if Bytes.length comp = 0 then Bytes.empty else g2_decompress comp

(* Device residency of a QAP.t (QAP.ml:11-16): rows in increasing Var order, coefficients lowest
   degree first, zero padded to n = degree target; uploaded once per value. *)
module Device (C : module type of Curve.Bls12_381) = struct
  open C
  module QAP = QAP.Make (Fr)
  module Poly = Fr.Poly

  let keys (q : QAP.t) = List.map fst (Var.Map.bindings q.QAP.v)

  let pad n p =
    let rec go i = function
      | _ when i = n -> []
      | [] -> Fr.zero :: go (i + 1) []
      | c :: cs -> c :: go (i + 1) cs
    in
    fr_bytes (go 0 p)

  let upload_qap (q : QAP.t) : int64 =
    let n = Poly.degree q.QAP.target in
    let flat m = cat (List.map (fun (_, p) -> pad n p) (Var.Map.bindings m)) in
    qap_load (flat q.QAP.v) (flat q.QAP.w) (flat q.QAP.y) (pad (n + 1) q.QAP.target) (List.length (keys q)) n

  (* the dense QAP stays on the device for as long as the [QAP.t] is in use (uploaded once) *)
  let qaps : QAP.t Resident.t = Resident.create ~capacity:4 qap_free
  let qap (q : QAP.t) : int64 = Resident.find_or_load qaps q upload_qap

  let index_of ks k =
    let rec go i = function [] -> assert false | x :: xs -> if x = k then i else go (i + 1) xs in
    go 0 ks

  let solution (q : QAP.t) sol = fr_bytes (List.map (fun k -> Var.Infix.(sol #! k)) (keys q))
end

(* Groth16.Make (C).prove, groth16.ml:235-237, with the body moved to the GPU.  keygen / verify are
   the reference's own (include Groth16.Make (C)); the result still satisfies Protocol.S. *)
module Groth16 (C : module type of Curve.Bls12_381) = struct
  include Groth16.Make (C)
  module D = Device (C)

  (* field order of zk_groth16_pkey: a, b1, d1, b2, d2, ti1, ti2, tiztd, ltd_mid *)
  let upload_key (pk : pkey) (q : qap) : int64 =
    let j = yojson_of_pkey pk in
    let ks = D.keys q in
    let n = C.Fr.Poly.degree q.D.QAP.target in
    let mids = Json.bindings j "ltd_mid" in
    let ti1 = g1_raw (Json.points j "ti1") and ti2 = g2_raw (Json.points j "ti2") in   (* n + 2 powers: the first n are used *)
    groth16_pk_load n (List.length ks)
      (Array.of_list (List.map (fun (k, _) -> D.index_of ks k) mids))
      [| g1_raw (Json.point j "a"); g1_raw (Json.point j "b1"); g1_raw (Json.point j "d1");
         g2_raw (Json.point j "b2"); g2_raw (Json.point j "d2");
         Bytes.sub ti1 0 (n * 96); Bytes.sub ti2 0 (n * 192); g1_raw (Json.points j "tiztd");
         g1_raw (cat (List.map snd mids)) |]
      (0, 1)

  (* the proving key is uploaded (and its window tables precomputed) on first use and stays
     resident next to [pk]: every later [prove] with the same key is one C call *)
  let keys : pkey Resident.t = Resident.create ~capacity:4 key_free
  let device_key (pk : pkey) (q : qap) : int64 = Resident.find_or_load keys pk (fun pk -> upload_key pk q)
  let release () = Resident.clear keys; Resident.clear D.qaps

  let prove rng (q : qap) (pk : pkey) sol : proof =
    let r = C.Fr.gen rng in                                (* groth16.ml:124 — r first *)
    let s = C.Fr.gen rng in                                (* groth16.ml:125 *)
    let out = groth16_prove (device_key pk q) (D.qap q) (D.solution q sol) (L.Fr.to_bytes r) (L.Fr.to_bytes s) in
    (* out = a | b | c, each uncompressed then compressed (96 + 48, 192 + 96, 96 + 48 bytes) *)
    let comp o len = Json.of_point (Bytes.sub out o len) in
    proof_of_yojson (`Assoc [ ("a", comp 96 48); ("b", comp 336 96); ("c", comp 528 48) ])

  (* groth16.ml:163-173 with the right-hand pairings moved to the left: one product, one final
     exponentiation.  [ab] must be [ab_b200 pk] (the reference's vkey.ab is a blst GT.t and cannot be
     compared with a device GT value). *)
  let ab_b200 (pk : pkey) : Gt_b200.t =
    let j = yojson_of_pkey pk in
    Gt_b200.product_bytes [ (g1_raw (Json.point j "a"), g2_raw (Json.point j "b2"), false) ]

  let verify_b200 ~(ab : Gt_b200.t) (w_io : C.Fr.t Var.Map.t) (vk : vkey) (pr : proof) =
    let jp = yojson_of_proof pr and jv = yojson_of_vkey vk in
    let io = Json.bindings jv "ltgm_io" in
    (* G1.dot vk.ltgm_io w_io *)
    let lin =
      Bytes.sub
        (g1_msm (g1_raw (cat (List.map snd io))) (fr_bytes (List.map (fun (k, _) -> Var.Infix.(w_io #! k)) io)))
        0 96
    in
    Gt_b200.equal ab
      (Gt_b200.product_bytes
         [ (g1_raw (Json.point jp "a"), g2_raw (Json.point jp "b"), false);
           (lin, g2_raw (Json.point jv "gm"), true);
           (g1_raw (Json.point jp "c"), g2_raw (Json.point jv "d"), true) ])
end

(* Pinocchio.Make (C).{NonZK, ZK}.prove, pinocchio.ml:536-538 / 559-561. *)
module Pinocchio (C : module type of Curve.Bls12_381) = struct
  module P = Pinocchio.Make (C)
  module D = Device (C)

  (* field order of zk_pinocchio_pkey: vv yy vav yay bvwy ww waw si v_all w_all one vt yt vavt yayt
     vbt wbt ybt wt wawt; j = yojson_of_pkey (pinocchio.ml:37-60) *)
  let upload_key (j : Json.t) (q : D.QAP.t) : int64 =
    let ks = D.keys q in
    let n = C.Fr.Poly.degree q.D.QAP.target in
    let mid = List.map fst (Json.bindings j "vv") in
    let m1 name = g1_raw (cat (List.map snd (Json.bindings j name)))
    and m2 name = g2_raw (cat (List.map snd (Json.bindings j name)))
    and p1 name = g1_raw (Json.point j name)
    and p2 name = g2_raw (Json.point j name) in
    let si = g1_raw (Json.points j "si") in
    pinocchio_pk_load n (List.length ks) (Array.of_list (List.map (D.index_of ks) mid))
      [| m1 "vv"; m1 "yy"; m1 "vav"; m1 "yay"; m1 "bvwy"; m2 "ww"; m2 "waw";
         Bytes.sub si 0 ((n + 1) * 96); m1 "v_all"; m1 "w_all";
         L.G1.to_bytes L.G1.one; p1 "vt"; p1 "yt"; p1 "vavt"; p1 "yayt";
         p1 "vbt"; p1 "wbt"; p1 "ybt"; p2 "wt"; p2 "wawt" |]
      (0, 1)

  (* results: vv | ww | yy | h | vavv | waww | yayy | bvwy, 144- and 288-byte point results, the
     compressed half last in each; the proof record of pinocchio.ml:195-208 as its yojson form *)
  let proof_json (out : bytes) : Json.t =
    let g1 o = Json.of_point (Bytes.sub out (o + 96) 48) and g2 o = Json.of_point (Bytes.sub out (o + 192) 96) in
    `Assoc [ ("vv", g1 0); ("ww", g2 144); ("yy", g1 432); ("h", g1 576); ("vavv", g1 720); ("waww", g2 864);
             ("yayy", g1 1152); ("bvwy", g1 1296) ]

  let prove_json hk (q : D.QAP.t) sol (d : bytes) : Json.t =
    proof_json (pinocchio_prove hk (D.qap q) (D.solution q sol) d)

  module NonZK = struct
    include P.NonZK
    let keys : pkey Resident.t = Resident.create ~capacity:4 key_free
    let prove _rng (q : qap) (pk : pkey) sol : proof =
      let hk = Resident.find_or_load keys pk (fun pk -> upload_key (yojson_of_pkey pk) q) in
      proof_of_yojson (prove_json hk q sol Bytes.empty)
  end

  module ZK = struct
    include P.ZK
    let keys : pkey Resident.t = Resident.create ~capacity:4 key_free
    let prove rng (q : qap) (pk : pkey) sol : proof =
      let dv = C.Fr.gen rng in                             (* pinocchio.ml:428 *)
      let dw = C.Fr.gen rng in                             (* :429 *)
      let dy = C.Fr.gen rng in                             (* :430 *)
      let hk = Resident.find_or_load keys pk (fun pk -> upload_key (yojson_of_pkey pk) q) in
      proof_of_yojson (prove_json hk q sol (fr_bytes [ dv; dw; dy ]))
  end

  let release () = Resident.clear NonZK.keys; Resident.clear ZK.keys; Resident.clear D.qaps
end

(* README.md:36-40 spellings.  The code has [Curve.Bls12_381] (curve.mli:56) and [Test.Make] /
   [Test.Make_suites] in library `test` (src/lib/test/test.mli:4-25); [Protocol] holds only the
   signature (protocol.ml:1-21).  These aliases make [Ecp.Bls12_381] and [Protocol.Test (F) (P)]
   resolve too. *)
module Ecp = struct module Bls12_381 = Bls12_381 end
module Protocol = struct
  include Protocol
  module Test_suites = Test.Make_suites     (* before [Test] below shadows the library *)
  module Test = Test.Make
end

(* The reference's own acceptance run (src/groth16/test/main.ml) on the GPU prover:
     open Zk
     module C = Curve.Bls12_381
     module G = Zkb200.Groth16 (C)                       (* : Protocol.S, prove on the GPU *)
     let () = Zkb200.init_devices [| 0; 1; 2; 3; 4; 5; 6; 7 |]
     module T = Test.Make_suites (C.Fr) (G)                                                  *)
